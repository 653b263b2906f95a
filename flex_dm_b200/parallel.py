"""Document-sharded data parallelism of the MFP step.  The reference has no distributed code
(train.py:25 is a commented-out MirroredStrategy stub); this is the one strategy the path needs.

Documents are independent through forward and backward (attention never crosses documents) and the loss is a batch
mean of per-document sums (metrics.py:265-277), so:

* every rank takes a contiguous slice of the global batch (``shard_documents``),
* computes gradients of ``(1 / B_global) * sum over its documents`` (``inv_batch`` of ``mfp_loss``),
* ``all_reduce(sum)`` of the flat fp32 gradient buffer gives the global-batch gradient: in one piece after the backward pass
  (``all_reduce_gradients``, the default of the train step) or one contiguous slice per backward stage started as soon as that
  stage's gradients are final (``all_reduce_gradient_slice``, ``MFP.enable_data_parallel(..., overlap=True)``),
* L2, per-variable clipnorm and Adam run after the reduce on every rank (replicated state), exactly the
  single-process semantics because clipping sees the reduced gradient,
* loss / score numerators / denominators are additive, so metric rows are all-reduced too (``reduce_metric_rows``);
  the L2 column is a function of the replicated weights and is not summed.

Everything here is backend-agnostic ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests).
"""
from typing import Dict, Tuple

import torch


def shard_bounds(batch_size: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced document ranges: the first ``batch_size % world_size`` ranks take one extra document."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(batch_size, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_documents(batch: Dict, rank: int, world_size: int) -> Dict:
    """This rank's documents of a global batch (every column is batch-major)."""
    some = next(iter(batch.values()))
    lo, hi = shard_bounds(int(some.shape[0]), rank, world_size)
    return {k: v[lo:hi] for k, v in batch.items()}


def all_reduce_gradients(dist, flat_grads: torch.Tensor) -> torch.Tensor:
    """One collective per step over the flat gradient buffer (11.25 MB for crello): sum of the ranks' partial gradients."""
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def all_reduce_gradient_slice(dist, flat_grads: torch.Tensor, lo: int, hi: int):
    """Start the all-reduce of one finished slice of the flat gradient buffer (a backward stage's variables) without waiting: on
    NCCL it runs on the communicator's stream under the backward of the layers below.  Returns the work handle; ``wait()`` it (a
    stream dependency, not a host block) before the optimiser reads the buffer."""
    if hi <= lo:
        return None
    return dist.all_reduce(flat_grads[lo:hi], op=dist.ReduceOp.SUM, async_op=True)


def reduce_metric_rows(dist, rows: torch.Tensor) -> torch.Tensor:
    """Sum per-rank metric rows ``[steps, 3F+2]`` (loss, score numerator / denominator per field, data loss); keep the
    replicated L2 column (last) as is."""
    rows = rows.clone()
    l2 = rows[:, -1].clone()
    dist.all_reduce(rows, op=dist.ReduceOp.SUM)
    rows[:, -1] = l2
    return rows


def broadcast_parameters(dist, flat_params: torch.Tensor, src: int = 0) -> torch.Tensor:
    dist.broadcast(flat_params, src)
    return flat_params
