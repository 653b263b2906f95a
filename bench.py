#!/usr/bin/env python
"""MFP train-step throughput on synthetic crello-shaped batches (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one full pass of the hot path over one batch: sample tasks -> mask/corrupt -> encoder -> L blocks ->
heads -> loss -> backward -> [NCCL all-reduce of the flat gradients] -> L2 + per-variable clip + Adam.
Workload at N GPUs: BASELINE.json configs[1] (crello Ours-IMP, masking_method=random, L=4, D=256, H=8, S=128,
256 documents per GPU, every document full length) -- weak scaling over documents.
Prints ONE JSON line on rank 0 (see the key list in DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MFP train-step elements/sec (crello seq_len=128)"
UNIT = "elements/s"
B_PER_GPU, SEQ_LEN, NUM_BLOCKS, LATENT = 256, 128, 4, 256
N_DEVICE_BATCHES = 4  # rotate distinct resident batches: 4 x 135 MB of inputs > the 126 MB L2

# BASELINE.json configs[0..4] = cfg1..cfg5.  B is documents per GPU, except cfg5 where it is the GLOBAL batch (strong scaling).
CONFIGS = {
    1: dict(name="cfg1: crello MFP --masking_method random, 2-layer d=256 h=8 seq_len=32 bs=4 (the reference's CPU-runnable case)",
            dataset="crello", method="random", L=2, S=32, B=4, scaling="weak"),
    2: dict(name="cfg2: crello Ours-IMP (--masking_method random) full config, seq_len=128", dataset="crello", method="random", L=4, S=128, B=256,
            scaling="weak"),
    3: dict(name="cfg3: crello Ours-EXP (--masking_method elem_pos_attr_img_txt) full config, seq_len=128", dataset="crello",
            method="elem_pos_attr_img_txt", L=4, S=128, B=256, scaling="weak"),
    4: dict(name="cfg4: rico Ours-EXP (--masking_method elem_pos_attr, sort_pos loss branch) full config, seq_len=128", dataset="rico",
            method="elem_pos_attr", L=4, S=128, B=256, scaling="weak"),
    5: dict(name="cfg5: crello Ours-IMP global bs=512 seq_len=128, documents sharded over the GPUs + NCCL gradient all-reduce", dataset="crello",
            method="random", L=4, S=128, B=512, scaling="strong"),
}


def flops_per_element(cols, S, L, D=256):
    """F_fwd = 2 D sum(d_num) + L (16 D^2 + 4 S D) + 2 D sum(W); train = 3 x."""
    from flex_dm_b200.spec import get_valid_input_columns

    d_num = sum(c["shape"][-1] for c in get_valid_input_columns(cols).values() if c["type"] == "numerical")
    w = sum(c["shape"][-1] * c["input_dim"] if c["type"] == "categorical" else c["shape"][-1] for c in get_valid_input_columns(cols).values())
    gemm_fwd = 2 * D * d_num + L * 16 * D * D + 2 * D * w
    attn_fwd = L * 4 * S * D
    return 3 * (gemm_fwd + attn_fwd), 3 * gemm_fwd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line).  The sampler is started before the
    warm-up steps (nvidia-smi needs ~0.1 s to come up, a 20-step timed region lasts 0.05 s) and every sample carries its arrival time;
    stop() keeps the samples that arrived between mark_begin() and mark_end().  When the region was shorter than one sampling period
    it falls back to the samples of the quarter second before its end -- the warm-up steps, the same kernels back to back -- and says so."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t1 = self.t1 if self.t1 is not None else time.perf_counter()
        t0 = self.t0 if self.t0 is not None else t1 - 1.0
        lines = [line for ts, line in self.lines if t0 <= ts <= t1]
        window = "timed region"
        if not lines:
            lines = [line for ts, line in self.lines if t1 - 0.25 <= ts <= t1 + 0.06]
            window = "warm-up + timed region (the timed region was shorter than one 50 ms sampling period)"
        sm, smax, reasons = [], [], set()
        for line in lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def cpu_cfg1_throughput(steps=10):
    """BASELINE.json configs[0] exactly (L=2, D=256, H=8, S=32, B=4, crello, random): the oracle's fp32 port on all host threads."""
    import torch

    from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
    from oracle import mfp_oracle as O

    w = CONFIGS[1]
    cols = make_input_columns(w["dataset"], max_length=50)
    batch = make_synthetic_batch(cols, w["B"], w["S"], seed=0, lengths="full")
    o = O.OracleMFP(cols, num_blocks=w["L"], masking_method=w["method"], dropout=0.1, l2=1e-2, dtype=torch.float32)
    for i in range(3):
        o.train_step(batch, seed=0, step=i)
    t0 = time.perf_counter()
    for i in range(steps):
        r = o.train_step(batch, seed=0, step=3 + i)
    el = time.perf_counter() - t0
    return {"value": w["B"] * w["S"] * steps / el, "unit": UNIT, "ms_per_step": 1e3 * el / steps, "steps": steps, "loss_last_step": r["loss"],
            "workload": w["name"]}


def cpu_port_throughput(budget_s=20.0, docs=8, threads=None):
    """The oracle's fp32 'port' of the reference train step (all masking variants, eager op sequence) on the host
    cores, on a bounded sample of the workload: `docs` full-length documents of the same schema / depth."""
    import torch

    from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
    from oracle import mfp_oracle as O

    torch.set_num_threads(threads or os.cpu_count() or 1)
    cols = make_input_columns("crello", max_length=SEQ_LEN)
    batch = make_synthetic_batch(cols, docs, SEQ_LEN, seed=0, lengths="full")
    o = O.OracleMFP(cols, num_blocks=NUM_BLOCKS, masking_method="random", dropout=0.1, l2=1e-2, dtype=torch.float32)
    o.train_step(batch, seed=0, step=0)  # warm-up
    steps, t0 = 0, time.perf_counter()
    while True:
        o.train_step(batch, seed=0, step=steps + 1)
        steps += 1
        el = time.perf_counter() - t0
        if el > budget_s or steps >= 50:
            break
    return {"value": docs * SEQ_LEN * steps / el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d steps of %d full-length crello documents (S=%d, L=%d, D=%d), fp32 PyTorch-CPU restatement of the TF eager op sequence "
                      "(TensorFlow is not installable here)" % (steps, docs, SEQ_LEN, NUM_BLOCKS, LATENT),
            "cfg1": cpu_cfg1_throughput()}


def run_reference(args):
    """--impl reference: the reference's own CPU path.  TF 2.8 cannot be installed (no wheel, Python 3.12), so this is
    the oracle port on all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
    from oracle import mfp_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: take every host core back
    w = CONFIGS[args.config]
    docs = min(8, w["B"])  # a bounded sample of the workload per step: 8 full-length documents (cfg1: its 4 documents = the whole config)
    S, L = w["S"], w["L"]
    cols = make_input_columns(w["dataset"], max_length=max(50, S))
    batch = make_synthetic_batch(cols, docs, S, seed=0, lengths="full")
    o = O.OracleMFP(cols, num_blocks=L, masking_method=w["method"], dropout=0.1, l2=1e-2, dtype=torch.float32)
    for i in range(args.warmup):
        o.train_step(batch, seed=0, step=i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        o.train_step(batch, seed=0, step=args.warmup + i)
    el = time.perf_counter() - t0
    value = docs * S * args.steps / el
    sample = "each step = %d full-length documents (S=%d, L=%d) of %s; fp32 PyTorch-CPU port of the reference op sequence" % (docs, S, L, w["name"])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s, CPU sample of %d documents per step" % (w["name"], docs)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def tfrecord_leg(model, timed, steps):
    """Trains from a TFRecord export of one synthetic batch (256 full-length crello documents, reshuffled every pass): every step the
    native reader parses 256 SequenceExamples on the host threads into pinned memory -- in the packed column format (55 MB of columns:
    the embedding rows no element carries are never written), with the dense format (135 MB) timed beside it."""
    import shutil
    import tempfile
    import time

    import torch

    from flex_dm_b200.data import DevicePrefetcher
    from flex_dm_b200.dataspec import DataSpec
    from flex_dm_b200.synthetic import write_synthetic_dataset

    root = tempfile.mkdtemp(prefix="flexdm_tfrecord_")
    try:
        write_synthetic_dataset(root, "crello", {"train": B_PER_GPU}, seq_len=SEQ_LEN, lengths="full", shards=2, seed=123)
        spec = DataSpec(os.path.join(root, "crello-spec.yml"), root, batch_size=B_PER_GPU)
        # host-only rate of the parser (no GPU work in flight), both batch formats
        n_host = 10
        host_ms = {}
        for packed in (False, True):
            it = iter(spec.make_dataset("train", shuffle=True, repeat=True, prefetch=0, pad_to=SEQ_LEN, packed=packed))
            for _ in range(3):
                next(it)
            t0 = time.perf_counter()
            for _ in range(n_host):
                next(it)
            host_ms[packed] = (time.perf_counter() - t0) * 1e3 / n_host
        row_host = torch.empty((model.engine.metrics_width,), dtype=torch.float32).pin_memory()

        def streamed(packed):
            feeder = DevicePrefetcher(model, iter(spec.make_dataset("train", shuffle=True, repeat=True, prefetch=3, pad_to=SEQ_LEN, packed=packed)))

            def step(i):
                row_host.copy_(model.train_step(next(feeder), staged=True), non_blocking=True)

            for i in range(3):
                step(i)
            return timed(step, steps)

        ms_dense = streamed(False)
        ms = streamed(True)
        # the same split resident in HBM (make_dataset(cache="device")): batches are gathered on the GPU, nothing is parsed or copied per step
        cached = spec.make_dataset("train", shuffle=True, repeat=True, cache="device", pad_to=SEQ_LEN)
        cached_it = iter(cached)

        def step_cached(i):
            row_host.copy_(model.train_step(next(cached_it), staged=True), non_blocking=True)

        for i in range(3):
            step_cached(i)
        ms_cached = timed(step_cached, steps)
        # the gather alone (kernel + the small index / context-column operations around it), against the HBM roofline: it reads and writes
        # every column of the batch once
        order = list(range(B_PER_GPU))
        for _ in range(3):
            cached.batch(order)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        n_gather = 20
        for k in range(n_gather):
            order = order[7:] + order[:7]
            got = cached.batch(order)
        ev1.record()
        torch.cuda.synchronize()
        gather_ms = ev0.elapsed_time(ev1) / n_gather
        gather_bytes = 2 * sum(t.numel() * t.element_size() for k, t in got.items() if t.dim() == 3)
        peak = 6650.0
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
        except Exception:
            pass
        return {"value": B_PER_GPU * SEQ_LEN * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
                "device_cached": {"value": B_PER_GPU * SEQ_LEN * steps / (ms_cached * 1e-3), "unit": UNIT, "ms_per_step": ms_cached / steps,
                                  "resident_bytes": cached.nbytes(),
                                  "gather": {"ms": gather_ms, "algorithmic_bytes": gather_bytes, "achieved": gather_bytes / (gather_ms * 1e-3) / 1e9,
                                             "unit": "GB/s", "peak": peak, "frac": gather_bytes / (gather_ms * 1e-3) / 1e9 / peak},
                                  "source": "the parsed split kept ragged in HBM, batches cut out by mfp_gather_documents (DataSpec.make_dataset(cache='device'))"},
                "dense_columns": {"value": B_PER_GPU * SEQ_LEN * steps / (ms_dense * 1e-3), "unit": UNIT, "ms_per_step": ms_dense / steps,
                                  "host_parse_ms_per_batch": host_ms[False]},
                "host_parse_ms_per_batch": host_ms[True], "host_threads": spec._threads,
                "source": "TFRecord shards of tf.train.SequenceExample (256 synthetic crello documents, S=128) -> DataSpec.make_dataset(shuffle, repeat, "
                          "prefetch=3, packed=True) -> libflexdm_io fdio_parse_batch_packed into pinned memory (packed numerical columns) -> DevicePrefetcher "
                          "-> MFP.train_step; dense_columns = the same with fdio_parse_batch (dense DataSpec.parse_fn layout)"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def input_columns_for(w):
    from flex_dm_b200.spec import make_input_columns

    return make_input_columns(w["dataset"], max_length=max(50, w["S"]))


def step0_golden(config):
    """The float64 oracle's loss of this configuration's first step (tools/make_bench_golden.py -> tests/golden/bench_step0.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "bench_step0.json"))).get("cfg%d" % config)
    except Exception:
        return None


class Workload:
    """One bench configuration on this rank: model, resident + pinned batches, the step function."""

    gemm_impl = 0
    docs_override = 0
    transport = "auto"  # gradient exchange at N > 1: "auto" (NVLS kernel when available), "nvls", "nccl"

    def __init__(self, config, world, rank, dev, dist):
        import torch

        from flex_dm_b200.mfp import MFP, Adam
        from flex_dm_b200.spec import make_synthetic_batch

        self.w = w = CONFIGS[config]
        self.config, self.world, self.rank, self.dev, self.dist = config, world, rank, dev, dist
        self.S, self.L = w["S"], w["L"]
        if w["scaling"] == "strong":
            if w["B"] % world:
                raise SystemExit("cfg%d: global batch %d does not divide over %d GPUs" % (config, w["B"], world))
            self.B = w["B"] // world
        else:
            self.B = w["B"]
        if Workload.docs_override:
            self.B = Workload.docs_override
        self.cols = cols = input_columns_for(w)
        self.model = model = MFP(cols, num_blocks=self.L, masking_method=w["method"], latent_dim=LATENT, dropout=0.1, l2=1e-2, seed=0, device=dev)
        model.compile(optimizer=Adam(learning_rate=1e-4, clipnorm=1.0))
        model.engine.set_gemm_impl(Workload.gemm_impl)
        if world > 1:
            model.enable_data_parallel(dist, world, transport=Workload.transport)
        # synthetic data: distinct batches per rank (seed = 1000 * rank + i), pinned on the host for the e2e leg
        host = [make_synthetic_batch(cols, self.B, self.S, seed=1000 * rank + i, lengths="full") for i in range(N_DEVICE_BATCHES)]
        needed = [k for k, c in model.input_columns.items() if k == "length" or c["is_sequence"]]
        from flex_dm_b200.data import pack_batch

        self.pinned = [{k: torch.from_numpy(b[k]).pin_memory() for k in needed} for b in host]
        # the input pipeline's packed column format (flex_dm_b200.data.pack_batch): embedding rows of elements whose type does not carry
        # the field are not stored -- the reference's filter_padding overwrites them with <UNUSED> before the model reads them
        self.pinned_packed = [{k: torch.from_numpy(v).pin_memory() for k, v in pack_batch({k: b[k] for k in needed}, cols).items()} for b in host]
        self.resident = [model.stage(b) for b in self.pinned_packed]  # HBM-resident batches, in the same (packed) format the pipeline delivers
        torch.cuda.synchronize()
        self.elements_per_step = self.B * self.S  # every document is full length: valid elements = B * S
        self.h2d_bytes_dense = sum(t.numel() * t.element_size() for t in self.pinned[0].values())
        self.h2d_bytes = sum(t.numel() * t.element_size() for b in self.pinned_packed for t in b.values()) // len(self.pinned_packed)

    def step_resident(self, i):
        return self.model.train_step(self.resident[i % N_DEVICE_BATCHES], staged=True)

    def global_loss(self, row):
        """Loss of the GLOBAL batch from this rank's metrics row: the additive columns are summed over the ranks (metrics.py:265-277)."""
        import torch

        from flex_dm_b200.parallel import reduce_metric_rows

        rows = row.detach().reshape(1, -1).to(self.dev)
        if self.world > 1:
            rows = reduce_metric_rows(self.dist, rows)
        return self.model.metrics_from_row(rows[0])["loss"]


def dp_check(config, world, rank, dev, dist, steps=3):
    """N GPUs on document shards vs ONE GPU on the concatenated batch (the reference is single-process: train.py:25), both with
    fixed-order reductions: per-step global loss, the weights after `steps` Adam updates, and a cross-rank checksum of the replicated
    parameters.  Ragged documents; shard r holds documents [r * B, (r + 1) * B) of the global batch."""
    import torch

    from flex_dm_b200.mfp import MFP, Adam
    from flex_dm_b200.parallel import reduce_metric_rows
    from flex_dm_b200.spec import make_synthetic_batch

    w = CONFIGS[config]
    B = w["B"] // world if w["scaling"] == "strong" else w["B"]
    S, L = w["S"], w["L"]
    cols = input_columns_for(w)

    def fresh():
        m = MFP(cols, num_blocks=L, masking_method=w["method"], latent_dim=LATENT, dropout=0.1, l2=1e-2, seed=0, device=dev)
        m.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
        m.set_deterministic(True)
        return m

    def shard(r, i):
        return make_synthetic_batch(cols, B, S, seed=7000 + 1000 * r + i, lengths="ragged")

    sharded = fresh()
    sharded.enable_data_parallel(dist, world, transport=Workload.transport)
    w0 = sharded.engine.params.clone()
    rows = torch.stack([sharded.train_step(shard(rank, i)).clone() for i in range(steps)])
    rows = reduce_metric_rows(dist, rows)
    losses = [sharded.metrics_from_row(r)["loss"] for r in rows]
    # replicated parameters: every rank must hold the same bits
    bits = sharded.engine.params.view(torch.int32).to(torch.int64)
    checksum = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
    gathered = [torch.zeros_like(checksum) for _ in range(world)]
    dist.all_gather(gathered, checksum)
    same = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    out = None
    if rank == 0:
        single = fresh()
        ref_losses = []
        for i in range(steps):
            parts = [shard(r, i) for r in range(world)]
            batch = {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
            ref_losses.append(single.metrics_from_row(single.train_step(batch))["loss"])
        torch.cuda.synchronize()
        # The key bias of every attention layer has an exactly-zero data gradient (softmax is shift-invariant): what both runs hold
        # there is rounding noise, which Adam normalises into +-lr moves.  Those 256-float slices are left out of the comparison.
        keep = torch.ones_like(w0)
        for name, (off, rows, vcols, ld, _) in single.engine.variables.items():
            if name.endswith("dense_key/bias"):
                keep[off:off + vcols] = 0.0
        upd_single, upd_sharded = (single.engine.params - w0) * keep, (sharded.engine.params - w0) * keep
        g_single, g_sharded = single.engine.grads * keep, sharded.engine.grads * keep
        out = {"steps": steps, "documents_per_gpu": B, "global_batch": B * world, "lengths": "ragged", "deterministic": True,
               "loss_single_gpu": ref_losses, "loss_sharded": losses,
               "max_rel_loss_diff": max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses)),
               "last_step_gradient_rel_l2_diff": ((g_single - g_sharded).norm() / g_single.norm()).item(),
               "weight_update_rel_l2_diff": ((upd_single - upd_sharded).norm() / upd_single.norm()).item(),
               "weight_update_max_abs": upd_single.abs().max().item(), "weight_update_max_abs_diff": (upd_single - upd_sharded).abs().max().item(),
               "note": "Adam moves every entry by ~lr per step whatever its size, so entries whose gradient is at rounding-noise level may differ by up to "
                       "2 lr per step between two correct runs; the L2 figures are the meaningful ones",
               "param_checksums_equal_across_ranks": same}
        del single
    del sharded
    torch.cuda.empty_cache()
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # ~0.6 s timed region: long enough for several nvidia-smi clock samples
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[config - 1]; the metric is quoted on 2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end and roofline legs (ncu launch-list runs)")
    ap.add_argument("--no-tfrecord", action="store_true", help="skip the TFRecord input-pipeline leg (N=1 only)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short value-only runs of the other BASELINE configs")
    ap.add_argument("--no-check-dp", action="store_true", help="N > 1: skip the sharded-vs-single-GPU equivalence check")
    ap.add_argument("--transport", default="auto", choices=["auto", "nvls", "nccl"], help="N > 1: gradient all-reduce transport")
    ap.add_argument("--docs-per-gpu", type=int, default=0, help="experiments: override the configuration's documents per GPU")
    ap.add_argument("--gemm-impl", type=int, default=0, choices=[0, 2], help="0 = TF32 product path (default), 2 = fp32-accurate 3xTF32 GEMMs + fp32 attention")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(steps):
            fn(i)
        stop.record()
        barrier()
        ms = torch.tensor([start.elapsed_time(stop)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)

    def measure_tf32_peak(n=8192, seconds=0.6):
        """Dense TF32 GEMM throughput of this GPU (cuBLAS through torch.matmul with TF32 allowed, fp32 operands in HBM, fp32 accumulate):
        the denominator for the tensor view of a path that computes in kind::tf32.  MEASURED_PEAKS.json holds only the bf16 figure."""
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            a = torch.randn((n, n), device=dev)
            b = torch.randn((n, n), device=dev)
            for _ in range(3):
                torch.matmul(a, b)
            torch.cuda.synchronize()
            best, total, iters = 0.0, 0.0, 0
            t_end = time.perf_counter() + seconds
            while time.perf_counter() < t_end:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(4):
                    torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 4
                best = max(best, 2.0 * n ** 3 / (ms * 1e-3) / 1e12)
                total += ms
                iters += 1
            return {"burst": best, "sustained": 2.0 * n ** 3 / (total / iters * 1e-3) / 1e12, "how": "torch.matmul fp32 %d^3 with allow_tf32, best of / mean over %d x 4 launches" % (n, iters)}
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old

    def value_leg(config, steps, warmup, before_timed=None, after_timed=None):
        """Device-resident throughput of one configuration: (workload, ms total, launches, step-0 global loss)."""
        wl = Workload(config, world, rank, dev, dist)
        loss0 = wl.global_loss(wl.step_resident(0))
        for i in range(1, warmup):
            wl.step_resident(i)
        launches0 = wl.model.engine.launch_count()
        if before_timed is not None:
            before_timed()
        ms = timed(wl.step_resident, steps)
        if after_timed is not None:
            after_timed()
        return wl, ms, wl.model.engine.launch_count() - launches0, loss0

    def check_step0(config, loss0):
        """The engine's first-step loss against the committed float64 oracle value (N = 1: rank 0's batch 0 is the golden batch)."""
        g = step0_golden(config)
        if g is None or world != 1 or Workload.docs_override:
            return None
        rel = abs(loss0 - g["loss"]) / abs(g["loss"])
        assert rel <= 2e-3, "cfg%d: step-0 loss %.6f differs from the oracle's %.6f (rel %.2e > 2e-3)" % (config, loss0, g["loss"], rel)
        return {"engine": loss0, "oracle_f64": g["loss"], "rel_err": rel, "tolerance": 2e-3, "source": "tests/golden/bench_step0.json"}

    Workload.gemm_impl = args.gemm_impl
    Workload.transport = args.transport
    Workload.docs_override = args.docs_per_gpu
    tf32 = measure_tf32_peak() if rank == 0 else None
    if world > 1:
        t = torch.tensor([tf32["sustained"] if tf32 else 0.0], device=dev)
        dist.broadcast(t, 0)
        tensor_peak = float(t.item())
    else:
        tensor_peak = tf32["sustained"]

    # ---- device-resident leg (value)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    wl, ms, launches, loss0 = value_leg(args.config, args.steps, args.warmup, before_timed=sampler.mark_begin if rank == 0 else None,
                                        after_timed=sampler.mark_end if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    model, w = wl.model, wl.w
    elements_per_step = wl.elements_per_step
    value = world * elements_per_step * args.steps / (ms * 1e-3)
    step0 = check_step0(args.config, loss0)
    h2d_bytes = wl.h2d_bytes
    d2h_bytes = model.engine.metrics_width * 4
    train_flops, gemm_flops = flops_per_element(wl.cols, wl.S, wl.L, LATENT)

    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms / args.steps, "gpu_launches": int(launches),
                              "note": "partial line (--no-e2e): not a bench result"}), flush=True)
        return

    # ---- end-to-end leg: host (pinned) batches through the public input pipeline (flex_dm_b200.data.DevicePrefetcher, what
    # MFP.fit uses): every step's columns are copied from pinned host memory inside the timed region, on a copy stream,
    # under the previous step's compute; the step's metrics row is read back to pinned host memory every step.
    from flex_dm_b200.data import DevicePrefetcher

    rows_host = torch.empty((args.steps, model.engine.metrics_width), dtype=torch.float32).pin_memory()

    def host_batches(source):
        i = 0
        while True:
            yield source[i % N_DEVICE_BATCHES]
            i += 1

    def e2e_leg(source):
        feeder = DevicePrefetcher(model, host_batches(source))

        def step_e2e(i):
            row = model.train_step(next(feeder), staged=True)
            rows_host[i % args.steps].copy_(row, non_blocking=True)

        for i in range(3):
            step_e2e(i)
        return timed(step_e2e, args.steps)

    ms_e2e = e2e_leg(wl.pinned_packed)       # packed columns (the pipeline's default format): the headline leg runs right after `value`
    ms_e2e_dense = e2e_leg(wl.pinned)        # dense DataSpec.parse_fn columns: 4136 B per element for crello
    e2e_value = world * elements_per_step * args.steps / (ms_e2e * 1e-3)
    last_loss = wl.global_loss(rows_host[args.steps - 1])
    assert np.isfinite(last_loss), last_loss

    # ---- input-side leg (N = 1, cfg2): the same step fed from TFRecord files through DataSpec.make_dataset (native SequenceExample parser on
    # host threads -> pinned batches -> DevicePrefetcher), i.e. train.py's own data path; reported beside e2e, not instead of it.
    input_pipeline = None
    if world == 1 and args.config == 2 and not args.no_tfrecord:
        try:
            input_pipeline = tfrecord_leg(model, timed, min(args.steps, 40))
        except Exception as e:  # an auxiliary leg must never cost the headline line (e.g. no writable temp directory on the box)
            input_pipeline = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- roofline of the dominant kernel (the TF32 tcgen05 GEMM): separate instrumented pass, CUDA events per launch
    prof_steps = 3
    torch.cuda.synchronize()
    model.engine.profile_begin()
    for i in range(prof_steps):
        wl.step_resident(i)
    prof = model.engine.profile_end()
    gemm_ms, gemm_launches, gemm_bytes = prof["gemm"]
    attn_ms, attn_launches, attn_bytes = prof["attention"]
    gemm_gbs = gemm_bytes / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else 0.0
    gemm_tflops = gemm_flops * elements_per_step * prof_steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # The dominant kernel class is the TF32 tcgen05 GEMM (all Dense forward / dgrad / wgrad contractions).  With fp32
    # activations and K = 256..512 these GEMMs sit below the ridge point, so the binding roofline is HBM: achieved =
    # algorithmic bytes of the step's GEMM launches (each operand and output once, counted by the engine) / their summed
    # CUDA-event time on the launching stream.  The tensor-pipe view of the same launches is given beside it.
    # measured DRAM bytes per launch of the same kernel class, from the committed ncu pass over one step of this workload
    # (profiles/kernel_traffic.json, written by tools/summarize_launches.py); null when that capture is absent
    traffic, traffic_src = None, None
    try:
        kt = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
        traffic = kt["classes"]["gemm"]["dram_bytes_per_launch"]
        traffic_src = "profiles/kernel_traffic.json: " + kt["source"]
    except Exception:
        pass
    step_s = ms / args.steps * 1e-3
    roofline = {"bound": "hbm", "kernel": "gemm_tf32_tcgen05", "achieved": gemm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gemm_gbs / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": gemm_bytes / max(gemm_launches, 1), "peak_source": src + " hbm_gbs",
                "gemm_ms_per_step": gemm_ms / prof_steps, "gemm_launches_per_step": gemm_launches // prof_steps,
                "gemm_algorithmic_bytes_per_step": gemm_bytes / prof_steps,
                "tensor": {"achieved": gemm_tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": gemm_tflops / tensor_peak,
                           "peak_source": "TF32 dense GEMM measured in this run (cuBLAS, %s): burst %.1f / sustained %.1f TFLOP/s; MEASURED_PEAKS.json "
                                          "bf16_tflops_sustained = %.1f for comparison" % (tf32["how"] if tf32 else "rank 0", tf32["burst"] if tf32 else 0.0, tensor_peak, bf16_peak),
                           "vs_bf16_peak_frac": gemm_tflops / bf16_peak},
                "attention": {"ms_per_step": attn_ms / prof_steps, "launches_per_step": attn_launches // prof_steps,
                              "achieved": attn_bytes / (attn_ms * 1e-3) / 1e9 if attn_ms > 0 else 0.0, "unit": "GB/s",
                              "frac": (attn_bytes / (attn_ms * 1e-3) / 1e9 / hbm_peak) if attn_ms > 0 else 0.0},
                "step_tensor_roofline_frac": train_flops * elements_per_step / step_s / 1e12 / tensor_peak,
                "step_tensor_roofline_frac_vs_bf16_peak": train_flops * elements_per_step / step_s / 1e12 / bf16_peak,
                # ~95 KB per element is the HBM floor of the step with per-sub-layer fusion and an fp32 residual stream
                "step_hbm_roofline_frac": 95e3 * elements_per_step / step_s / 1e9 / hbm_peak}

    # ---- N > 1: the sharded step against the single-process step on the concatenated batch
    check = None
    if world > 1 and not args.no_check_dp:
        check = dp_check(args.config, world, rank, dev, dist)

    # ---- the other BASELINE configs, value only (short runs; each line carries its own step-0 parity check and roofline fractions)
    others = None
    if not args.no_other_configs and args.config == 2:
        others = {}
        for c in (3, 4, 5):
            if CONFIGS[c]["scaling"] == "strong" and CONFIGS[c]["B"] % world:
                continue
            o_wl, o_ms, o_launches, o_loss0 = value_leg(c, min(args.steps, 100), 5)
            o_steps = min(args.steps, 100)
            tf, _ = flops_per_element(o_wl.cols, o_wl.S, o_wl.L, LATENT)
            o_val = world * o_wl.elements_per_step * o_steps / (o_ms * 1e-3)
            others["cfg%d" % c] = {"workload": "%s, %d documents per GPU" % (o_wl.w["name"], o_wl.B), "value": o_val, "unit": UNIT,
                                   "ms_per_step": o_ms / o_steps, "steps": o_steps, "scaling": o_wl.w["scaling"], "global_batch": o_wl.B * world,
                                   "gpu_launches_per_step": o_launches / o_steps, "train_flop_per_element": tf,
                                   "step_tensor_roofline_frac": tf * o_wl.elements_per_step / (o_ms / o_steps * 1e-3) / 1e12 / tensor_peak,
                                   "loss_step0": check_step0(c, o_loss0) or {"engine": o_loss0}}
            del o_wl
            torch.cuda.empty_cache()

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu = None if args.no_cpu_baseline else cpu_port_throughput(args.cpu_budget)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
                "dtype": "tf32" if args.gemm_impl == 0 else "f32 (3xTF32 tensor-core GEMMs, fp32 attention)",
                "data": "synthetic",
                "config": {"workload": "%s: L=%d D=256 H=8 FFN=512, %d documents per GPU, all documents full length, dropout=0.1 l2=1e-2 "
                                       "Adam(1e-4, clipnorm=1.0)" % (w["name"], wl.L, wl.B),
                           "global_batch": wl.B * world, "seq_len": wl.S, "parallelism": "dp%d" % world,
                           "gradient_exchange": None if world == 1 else ("one NVLS kernel on the step's stream (multimem.ld_reduce / multimem.st, csrc/allreduce.cu)"
                                                                         if model._nvls is not None else "ncclAllReduce of the flat gradient buffer"),
                           "l2_flush": "inputs larger than L2: %d distinct resident batches (packed columns, %.0f MB) rotate; activations per step 1.6 GB" % (N_DEVICE_BATCHES, N_DEVICE_BATCHES * h2d_bytes / 1e6),
                           "loss_step0": step0 if step0 is not None else {"engine": loss0, "note": "global loss (metric rows summed over ranks); the oracle value is pinned at N=1"},
                           "loss_last_step": last_loss},
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                        "h2d_gbs_per_gpu": h2d_bytes / (ms_e2e / args.steps * 1e-3) / 1e9,
                        "input_format": "packed columns in pinned host memory (flex_dm_b200.data.pack_batch: embedding rows of elements whose type does not "
                                        "carry the field are not stored) -> DevicePrefetcher (copy stream) -> MFP.train_step; metrics row read back every step",
                        "dense_columns": {"value": world * elements_per_step * args.steps / (ms_e2e_dense * 1e-3), "unit": UNIT,
                                          "ms_per_step": ms_e2e_dense / args.steps, "h2d_bytes_per_step": wl.h2d_bytes_dense,
                                          "h2d_gbs_per_gpu": wl.h2d_bytes_dense / (ms_e2e_dense / args.steps * 1e-3) / 1e9}},
                "gpu_launches": int(launches), "clocks": clocks}
        if check is not None:
            line["dp_check"] = check
        if others:
            line["other_configs"] = others
        if input_pipeline is not None:
            line["input_pipeline"] = input_pipeline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
