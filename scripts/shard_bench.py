#!/usr/bin/env python
"""Row-sharded training + ranking benchmark (BASELINE.json configs[3]/[4] shapes), run under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
      scripts/shard_bench.py cfg5 [--ent-per-gpu 1250000]

cfg5: ComplEx k=1000 eta=50, R=1000, B=8192 positives per GPU, entity table sharded by rows over the N
GPUs (1.25 M entities = 10 GB per GPU by default; 8 GPUs -> 10 M entities / 80 GB), lazy Adam (the dense
rule would stream 8 x 10 GB per step and shard), then full-entity ranking of 1024 test triples, both sides.
cfg4: RotatE k=200 eta=30, 123,182 entities, 37 relations, B=10,791 per GPU, dense Adam.
Timing: CUDA events, max over ranks; inputs resident in HBM."""
import argparse
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ampligraph_b200.parallel import ShardedTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("cfg", choices=["cfg4", "cfg5"])
ap.add_argument("--ent-per-gpu", type=int, default=1250000)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
if a.cfg == "cfg5":
    model, k, eta, E, R, B, lazy, loss = "ComplEx", 1000, 50, a.ent_per_gpu * world, 1000, 8192, True, "self_adversarial"
else:
    model, k, eta, E, R, B, lazy, loss = "RotatE", 200, 30, 123182, 37, 10791, False, "self_adversarial"
tr = ShardedTrainer(model, k, eta, E, R, local, lazy=lazy, loss=loss, optimizer="adam")
tr.eng.init_glorot_uniform(1 + rank)  # each shard its own stream
torch.cuda.synchronize()
tr.hdl.barrier(channel=0)
rng = np.random.default_rng(100 + rank)
nb = 4
batches = [torch.as_tensor(np.stack([rng.integers(0, E, B), rng.integers(0, R, B), rng.integers(0, E, B)], 1).astype(np.int32)).cuda()
           for _ in range(nb)]
for i in range(a.warmup):
    tr.train_step(batches[i % nb], None, seed=7 + rank, step=i)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    tr.train_step(batches[i % nb], None, seed=7 + rank, step=a.warmup + i)
e1.record()
torch.cuda.synchronize(); dist.barrier()
t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = t.item() / a.steps
# attribution of one step (separate pass, events between the phases)
evs = []
for i in range(3):
    evs = []
    tr.train_step(batches[i % nb], None, seed=7 + rank, step=1000 + i, events=evs)
torch.cuda.synchronize()
phases = [round(evs[j].elapsed_time(evs[j + 1]), 3) for j in range(4)]
ld = tr.eng.ld
alg = 2 * (3 + eta) * ld * 4 * B  # per rank per step
# ranking: 1024 test triples, both sides, against ALL entities (each rank its shard, counts summed)
q = batches[0][:1024].contiguous()
for side in ("s", "o"):
    tr.rank_counts(q, side)
torch.cuda.synchronize(); dist.barrier()
r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
r0.record()
for side in ("s", "o"):
    cnt = tr.rank_counts(q, side)
r1.record()
torch.cuda.synchronize()
tr_ms = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device="cuda")
dist.all_reduce(tr_ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print({"cfg": a.cfg, "n_gpus": world, "entities": E, "table_GB": round(E * ld * 4 / 1e9, 2), "ms_per_step": round(ms, 3),
           "triples_per_s": round(world * B * (1 + eta) / (ms / 1e3)), "alg_GBps_per_gpu": round(alg / (ms / 1e3) / 1e9, 1),
           "nvlink_rows_GB_per_gpu_per_step_each_way(one gather)": round((3 + eta) * B * ld * 4 * (world - 1) / world / 1e9, 3),
           "stash": os.environ.get("KGE_B200_STASH", "1"),
           "rank_1024x2sides_ms": round(tr_ms.item(), 2), "rank_Gscores_per_s": round(2 * 1024 * E / (tr_ms.item() / 1e3) / 1e9, 1),
           "rank0_phase_ms[kernel,barrier0,optimizers,barrier1]": phases, "G": None,
           "loss": tr.eng.read_loss()}, flush=True)
tr.close()
dist.destroy_process_group()
