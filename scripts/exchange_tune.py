#!/usr/bin/env python
"""Development aid (run under torchrun on N GPUs): times the data-parallel cfg2 step for several settings of the exchange
kernel in ONE process group -- mode (p2p / nvls) x KGE_B200_EXCHANGE_TRIPS -- and prints step time, train-kernel time and
the exchange kernel's phase stamps.  torchrun --nproc-per-node N scripts/exchange_tune.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ampligraph_b200.engine import KGEEngine  # noqa: E402
from ampligraph_b200.parallel import DataParallelTrainer, batch_slot  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
C = bench.CFG
B, eta = C["batch"], C["eta"]
data_np = bench.synthetic_kg(C["n_ent"], C["n_rel"], C["n_triples"])
data = torch.as_tensor(data_np).to(dev)
nb = len(data_np) // B
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def make_engine(alloc):
    return KGEEngine(C["model"], C["k"], eta, C["n_ent"], C["n_rel"], loss=C["loss"], loss_params=C["loss_params"],
                     optimizer="adam", optimizer_params={"learning_rate": C["lr"]}, device=local, table_alloc=alloc)


def measure(dp, steps=20, warmup=5):
    def step(i, ev=None):
        j = batch_slot(i, world, rank, nb)
        b = data[j * B:(j + 1) * B]
        if ev: ev[0].record()
        dp.train_step(b, None, seed=1234, step=i, kernel_done=ev[1] if ev else None)
        if ev: ev[2].record()
    for i in range(warmup):
        step(i)
    dist.barrier(); torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for i in range(steps):
        flush.fill_(i)
        step(warmup + i, evs[i])
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([sum(e[0].elapsed_time(e[2]) for e in evs) / steps, sum(e[0].elapsed_time(e[1]) for e in evs) / steps],
                     dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dp.trace_exchange(True)
    rows = []
    for i in range(7):
        flush.fill_(i)
        step(1000 + i)
        dist.barrier(); torch.cuda.synchronize()
        rows.append(dp.exchange_phases_us())
    dp.trace_exchange(False)
    med = {k: round(float(np.median([r[k] for r in rows])), 1) for k in rows[0]}
    return t[0].item(), t[1].item(), med


for mode in sys.argv[1:] or ["nvls", "p2p"]:
    try:
        dp = DataParallelTrainer(make_engine, mode=mode)
    except Exception as e:
        if rank == 0:
            print(mode, "unavailable:", repr(e), flush=True)
        continue
    dp.eng.init_glorot_uniform(3)
    dp.eng.set_hot_entities(triples=data_np)
    for trips in ([1, 2, 3, 4, 6] if mode == "nvls" else [1]):
        os.environ["KGE_B200_EXCHANGE_TRIPS"] = str(trips)
        ms, mk, ph = measure(dp)
        if rank == 0:
            print("N=%d mode=%-4s trips=%d  step %.4f ms  train kernel %.4f ms  tail %.1f us  %.3f G triples/s  phases %s"
                  % (world, mode, trips, ms, mk, 1e3 * (ms - mk), world * B * (1 + eta) / ms / 1e6, ph), flush=True)
    dp.close()
    del dp
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
