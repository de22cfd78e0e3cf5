"""The host path (HMC.step on a pinned host ensemble) of config 2 at N ranks against the raw PCIe ceiling:
every rank moves its shard of the 2^20-particle ensemble (D x P/N float32) host -> device and back, all ranks at once.

    python -m torch.distributed.run --nproc-per-node N profiles/e2e_probe_ranks.py

Separates the two possible causes of the flat e2e curve of the scaling run (SCALE_r01: 4.75e9 / 5.98e9 / 6.20e9 /
8.16e9 particle-leapfrog-steps/s at 1 / 2 / 4 / 8 GPUs): the library's chunking on small shards, or N processes
sharing the host's memory / PCIe root complexes.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402
import bench  # noqa: E402

KB = 1.380649e-23
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
D, P, L, h = 100, 1 << 20, 50, 0.05
Pl = P // world
ctx = E._lib.Context.get(local)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def maxover(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- raw copies: concurrent H2D + D2H of the shard on two streams, all ranks together ----
qh = torch.randn(D, Pl).pin_memory()
qh2 = torch.empty(D, Pl).pin_memory()
d1 = torch.empty(D, Pl, device=dev)
d2 = torch.empty(D, Pl, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mode in ("h2d+d2h", "h2d", "d2h"):
    barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        if mode != "d2h":
            with torch.cuda.stream(s1):
                d1.copy_(qh, non_blocking=True)
        if mode != "h2d":
            with torch.cuda.stream(s2):
                qh2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    ms = maxover((time.perf_counter() - t0) / 10 * 1e3)
    nbytes = D * Pl * 4 * (2 if mode == "h2d+d2h" else 1)
    if rank == 0:
        print(f"N={world} raw {mode:8s} of {D * Pl * 4 / 1e6:.1f} MB per rank: {ms:.3f} ms (max over ranks), "
              f"{nbytes / ms / 1e6:.1f} GB/s per rank, {world * nbytes / ms / 1e6:.1f} GB/s aggregate", flush=True)

# ---- the library's host path on the same shard ----
pot = E.GaussianPotential(precision=bench.make_precision(D))
ens = E.Ensemble(D, Pl, dtype=np.float32, seed=1, particleOffset=rank * Pl)
ens.q = qh.numpy()
ens.mass = torch.ones(Pl, dtype=torch.float32).pin_memory().numpy()
hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pot, rng="philox", seed=1, bugCompat=False)
acc = torch.empty(Pl, dtype=torch.uint8).pin_memory().numpy()
for mb in (2, 4, 8, 16, 32):
    ctx.set_option("host_chunk_mb", mb)
    hmc.step(1 / KB, accept=acc)
    barrier()
    t0 = time.perf_counter()
    for _ in range(6):
        hmc.step(1 / KB, accept=acc)
    torch.cuda.synchronize()
    ms = maxover((time.perf_counter() - t0) / 6 * 1e3)
    if rank == 0:
        print(f"N={world} HMC.step host path, host_chunk_mb={mb:3d}: {ms:.3f} ms / iteration (max over ranks), "
              f"{2 * D * Pl * 4 / ms / 1e6:.1f} GB/s per rank both directions, "
              f"{P * L / ms / 1e-3:.3e} particle-leapfrog-steps/s", flush=True)
# device-resident iteration of the shard for scale
qd = torch.randn(D, Pl, device=dev)
ens_d = E.Ensemble(D, Pl, dtype=np.float32, device=dev, seed=1, particleOffset=rank * Pl)
hmc_d = E.HMC(ens_d, L * h + 1e-9, h, None, potential=pot, seed=1, bugCompat=False)
for _ in range(3):
    hmc_d.step(1 / KB)
barrier()
t0 = time.perf_counter()
for _ in range(20):
    hmc_d.step(1 / KB)
torch.cuda.synchronize()
ms = maxover((time.perf_counter() - t0) / 20 * 1e3)
if rank == 0:
    print(f"N={world} device-resident iteration of the shard: {ms:.3f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
