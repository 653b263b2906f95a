import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# GPU run order: kernel-level parity first (GEMM / attention / optimiser units, the reference-run goldens, the benchmarked cfg2 shape),
# then the flows built on top of them (contexts, input pipeline, fit / train()) -- a failure in a flow must not hide a kernel result.
_GPU_ORDER = ["test_gpu_parity.py", "test_golden_reference.py", "test_gpu_cfg2_shape.py", "test_gpu_determinism.py", "test_gpu_context.py",
              "test_gpu_input_pipeline.py", "test_gpu_train_flow.py"]


def pytest_collection_modifyitems(config, items):
    def rank(item):
        name = os.path.basename(str(item.fspath))
        return _GPU_ORDER.index(name) if name in _GPU_ORDER else -1  # CPU-side files keep their place in front

    items.sort(key=rank)  # stable: order within a file is kept
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
