"""Config 2 through the host path (HMC.step on a pinned host ensemble): chunk-size sensitivity.
    python profiles/e2e_probe.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402
import bench  # noqa: E402

KB = 1.380649e-23
D, P, L, h = 100, 1 << 20, 50, 0.05
ctx = E._lib.Context.get(0)
pot = E.GaussianPotential(precision=bench.make_precision(D))
q = torch.randn(D, P).pin_memory()
ens = E.Ensemble(D, P, dtype=np.float32, seed=1)
ens.q = q.numpy()
ens.mass = torch.ones(P, dtype=torch.float32).pin_memory().numpy()
hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pot, rng="philox", seed=1, bugCompat=False)
acc = torch.empty(P, dtype=torch.uint8).pin_memory().numpy()
for mb in (4, 8, 16, 32, 64, 128):
    ctx.set_option("host_chunk_mb", mb)
    hmc.step(1 / KB, accept=acc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        hmc.step(1 / KB, accept=acc)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"host_chunk_mb={mb:4d}: {ms:.2f} ms / iteration, {2 * D * P * 4 / ms / 1e6:.1f} GB/s both directions")
# raw PCIe reference: concurrent H2D + D2H of the same volume on two streams
d = torch.empty(D, P, device="cuda")
d2 = torch.empty(D, P, device="cuda")
q2 = torch.empty(D, P).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(q, non_blocking=True)
    with torch.cuda.stream(s2):
        q2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 5 * 1e3
print(f"raw concurrent H2D + D2H of 2 x {D * P * 4 / 1e6:.0f} MB: {ms:.2f} ms, {2 * D * P * 4 / ms / 1e6:.1f} GB/s both directions")
