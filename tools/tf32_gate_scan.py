"""CPU tool: which ReLU gates of a golden fixture lie inside TF32's rounding error, and how much gradient hangs on each.

For every fixture of tests/test_golden_reference.py the float64 oracle is run as is, with its matrix products on TF32-rounded operands
(round-to-nearest and truncation, ``oracle.emulate_tf32``), and then once per gate that either rounding mode flips, with that single gate
forced to the other side.  Printed per gate: block, (document, oracle row, unit), the pre-activation in float64 / RNA / truncation, and the
largest per-variable change of the gradient (L2, relative to the variable's gradient norm) that flipping it causes.  A gate with a change
above ``tests/helpers.py::GRAD_REL_L2`` makes the fixture's gradient checks depend on the rounding mode of the product path: list it in
pick another seed for it (tests/golden/make_golden.py): round 1's ``--context canvas`` fixture was regenerated for that reason.
This is how the open item of ``--context canvas`` was traced.  Usage: python tools/tf32_gate_scan.py [case ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mfp_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_golden_reference import CASES, oracle_step, tf32_truncate  # noqa: E402


def worst_change(a, b):
    out = (0.0, "")
    for name in a:
        if name.endswith("dense_key/bias"):  # exact data gradient 0: nothing to be relative to
            continue
        d = float(np.linalg.norm(a[name] - b[name]) / max(np.linalg.norm(b[name]), 1e-30))
        out = max(out, (d, name))
    return out


def main(cases):
    for case in cases:
        pre, pre_rna, pre_rz = {}, {}, {}
        exact, _ = oracle_step(case, record=pre)
        rna, _ = oracle_step(case, tf32=O.tf32_round, record=pre_rna)
        rz, _ = oracle_step(case, tf32=tf32_truncate, record=pre_rz)
        print("%s: TF32 emulation vs float64, worst variable: RNA %.4f (%s), truncation %.4f (%s)" % ((case,) + worst_change(rna, exact) + worst_change(rz, exact)))
        for block in sorted(pre):
            a, b, c = pre[block], pre_rna[block], pre_rz[block]
            flipped = (((a > 0) != (b > 0)) | ((a > 0) != (c > 0))).nonzero()
            print("  block %d: pre-activation rms %.3f, max TF32 error %.1e (RNA) %.1e (truncation), %d gate(s) change side" %
                  (block, float(a.pow(2).mean().sqrt()), float((b - a).abs().max()), float((c - a).abs().max()), len(flipped)))
            for index in flipped:
                index = tuple(int(i) for i in index)
                if float(a[index]) > 0:  # open in float64: shut it there
                    changed, _ = oracle_step(case, closed_gates=[(block, index)])
                    d, name = worst_change(changed, exact)
                else:  # shut in float64, open under one of the roundings: shut it in that run
                    mode, base = (O.tf32_round, rna) if float(b[index]) > 0 else (tf32_truncate, rz)
                    changed, _ = oracle_step(case, closed_gates=[(block, index)], tf32=mode)
                    d, name = worst_change(changed, base)
                flag = "  <-- over GRAD_REL_L2" if d > H.GRAD_REL_L2 else ""
                print("    %s f64 %+.2e rna %+.2e rz %+.2e: flipping it moves %s by %.4f%s" %
                      (index, float(a[index]), float(b[index]), float(c[index]), name.replace("model/", ""), d, flag))


if __name__ == "__main__":
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    main(sys.argv[1:] or list(CASES))
