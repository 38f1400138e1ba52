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


class DataParallelTrainer:
    """Replicated tables, one rank per GPU, the reference's train_step on the GLOBAL batch.

    Each rank runs the fused kernel on its own slice of the batch into its own gradient table.
    Then, preferred path ("p2p"): tables and gradient tables live in torch symmetric memory, so
    every rank can address every peer's copy through NVLink; ONE kernel per table
    (kge_optimizer_step_sharded) reduce-scatters the gradients by peer loads, runs the optimizer
    on this rank's row shard (slots are sharded: 1/N of the Adam state per GPU) and all-gathers
    the new rows by peer stores.  Two stream-ordered cross-rank barriers bracket it.
    Fallback ("nccl"): SUM all-reduce of the gradient tables + the full optimizer on every replica.
    Both equal the single-GPU step on the concatenated batch up to fp32 summation order, because
    the loss is a SUM over the batch (loss_functions.py:214).
    """

    def __init__(self, make_engine, mode="auto", group=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.mode = "single" if self.world == 1 else mode
        self.hdl = None
        if self.mode in ("auto", "p2p"):
            try:
                self.eng = make_engine(self._symmetric_alloc)
                self._finish_p2p_setup()
                self.mode = "p2p"
            except Exception as e:  # no symmetric memory / no peer access: NCCL path
                if mode == "p2p":
                    raise
                self.p2p_error = repr(e)
                self.mode = "nccl"
                self.hdl = None
                self.eng = make_engine(None)
        else:
            self.eng = make_engine(None)

    # ---- symmetric-memory plumbing ------------------------------------------------
    def _symmetric_alloc(self, n_ent, n_rel, ld, device):
        """One symmetric buffer [ent | rel | g_ent | g_rel] so a single rendezvous maps all four."""
        import torch.distributed._symmetric_memory as symm_mem
        rows = 2 * (n_ent + n_rel)
        self._buf = symm_mem.empty((rows, ld), dtype=torch.float32, device=device)
        self._buf.zero_()
        pg = self.group if self.group is not None else dist.group.WORLD
        self.hdl = symm_mem.rendezvous(self._buf, pg.group_name)
        b = self._buf
        self._offsets = {"ent": 0, "rel": n_ent, "g_ent": n_ent + n_rel, "g_rel": 2 * n_ent + n_rel}
        return (b[0:n_ent], b[n_ent:n_ent + n_rel], b[n_ent + n_rel:2 * n_ent + n_rel], b[2 * n_ent + n_rel:rows])

    def _finish_p2p_setup(self):
        C, eng = self._C, self.eng
        ld = eng.ld
        base = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(base) == self.world and base[self.rank] == self._buf.data_ptr()
        mk = lambda key: (C.c_void_p * self.world)(*[b + self._offsets[key] * ld * 4 for b in base])
        self._ptrs = {k: mk(k) for k in ("ent", "rel", "g_ent", "g_rel")}
        self.shards = {"ent": row_shard(eng.n_ent, self.world, self.rank), "rel": row_shard(eng.n_rel, self.world, self.rank)}
        # optimizer slots only for this rank's row shards
        for key in ("ent", "rel"):
            lo, hi = self.shards[key]
            eng.slots[key] = [None if s is None else s[lo:hi].clone() for s in eng.slots[key]]
        torch.cuda.synchronize()
        self.hdl.barrier(channel=0)

    # ---- one global step ------------------------------------------------------------
    def train_step(self, batch, negatives=None, seed=0, step=0, kernel_done=None):
        eng = self.eng
        eng.forward_backward(batch, negatives, seed=seed, step=step)
        if kernel_done is not None:
            kernel_done.record()  # CUDA event: lets bench.py time the fused kernel alone
        if self.mode == "single":
            eng.apply_gradients()
        elif self.mode == "nccl":
            allreduce_sum_([eng.g_ent, eng.g_rel], self.group)
            eng.apply_gradients()
        else:
            C, _lib = self._C, self._lib
            eng.t += 1
            st = eng._stream()
            self.hdl.barrier(channel=0)  # every rank's gradient table is complete
            for key, rows in (("ent", eng.n_ent), ("rel", eng.n_rel)):
                lo, hi = self.shards[key]
                s0, s1 = eng.slots[key]
                _lib.check(eng.lib.kge_optimizer_step_sharded(
                    eng.h, C.byref(eng.opt_cfg), eng.t, self.world, self.rank, self._ptrs[key], self._ptrs["g_" + key],
                    C.c_void_p(s0.data_ptr() if s0 is not None else 0), C.c_void_p(s1.data_ptr() if s1 is not None else 0),
                    lo, hi, C.c_void_p(eng.loss_acc.data_ptr() + 8), st))
            self.hdl.barrier(channel=1)  # all parameters delivered, all peers done reading my gradients
            eng.g_ent.zero_()
            eng.g_rel.zero_()
            eng.launches += 4

    def close(self):
        self.eng.close()
