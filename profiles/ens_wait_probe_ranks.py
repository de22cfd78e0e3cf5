"""The fused ensemble run with the in-kernel all-reduce at 2^19 particles PER RANK (the 8-GPU shard of config 5) under
torchrun: per-iteration time against the single-process number, waits of the compute warps, master phase stamps.
    python -m torch.distributed.run --nproc-per-node N profiles/ens_wait_probe_ranks.py [iterations] [L]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import physicsbasedbayesianinference_b200 as E

KB = 1.380649e-23
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
D, L, h = 10, int(sys.argv[2]) if len(sys.argv) > 2 else 20, 0.05
P = 1 << 19
ctx = E._lib.Context.get()
ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1, particleOffset=rank * P)
ens.setPosition(1.0)
hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.FunnelPotential(D, 3.0), seed=1, bugCompat=False)
hmc.run(50, 1 / KB, adapt=True, keepNumSteps=True, adaptLag=2)
for lag, dbg in ((2, 0), (3, 0), (3, 24), (4, 0)):
    hmc.run(20, 1 / KB, adapt=False, keepNumSteps=True, adaptLag=lag)
    torch.cuda.synchronize()
    dist.barrier()
    if dbg:
        ctx.set_option("ens_debug", dbg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hmc.run(n, 1 / KB, adapt=False, keepNumSteps=True, adaptLag=lag)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} P/rank=2^19 L={L} lag={lag} debug={dbg}: {1e3 * t.item() / n:.2f} us/iteration (max over ranks)",
              file=sys.stderr, flush=True)
    if dbg and rank in (0, world - 1):
        ctx.set_option("ens_debug_dump", 12)
    if dbg:
        ctx.set_option("ens_debug", 0)
dist.destroy_process_group()
