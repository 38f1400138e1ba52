"""Multi-GPU host logic: one process per GPU (torch.distributed), data-parallel over positives.

The training path shards naturally over positives: every rank runs the fused
kernel on its own slice of the global batch into its own gradient tables, the tables are
summed with ONE collective (the loss is a SUM over the batch, loss_functions.py:214, so the
summed gradient equals the single-GPU gradient of the global batch), and every rank applies
the identical dense optimizer step to its replica.  Ranking shards over entity rows: each rank
counts against its row range and the int32 counts are summed (rank counts are additive over
entity partitions, ScoringBasedEmbeddingModel.py:1449-1452).
"""
import torch
import torch.distributed as dist


def batch_slot(step, world, rank, n_batches):
    """Which sequential batch of the epoch rank `rank` trains at global step `step`
    (rank-major interleave keeps the reference's sequential order across the job)."""
    return (step * world + rank) % n_batches


def row_shard(n_rows, world, rank):
    """Contiguous row range [begin, end) of rank `rank` (NVSwitch is uniform: no topology awareness)."""
    per = (n_rows + world - 1) // world
    lo = min(rank * per, n_rows)
    return lo, min(lo + per, n_rows)


def allreduce_sum_(tensors, group=None):
    """In-place SUM all-reduce of the gradient tables / rank counters (NCCL on GPUs, gloo in CPU tests)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors
