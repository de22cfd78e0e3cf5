"""Per-iteration cost of the small-D kernel (config 5: funnel, D = 10, P = 2^22): kernel only (with and
without the statistics partials) and through HMC.run (stale-statistics pipeline).
    python profiles/k1_probe.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402

KB = 1.380649e-23
D, P, h = 10, 1 << 22, 0.05
ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
ens.setPosition(1.0)
ctx = E._lib.Context.get()
for waves in (1, 2, 4, 8, 16, 100000):
    ctx.set_option("small_waves", waves)
    hmc = E.HMC(ens, 4 * h + 1e-9, h, None, potential=E.FunnelPotential(D, 3.0), seed=1, bugCompat=False)
    st = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
    for stats in (None, st):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(43):
            if i == 3:
                e0.record()
            hmc.step(1 / KB, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        print(f"waves={waves} L=4 stats={'on' if stats is not None else 'off'}: {e0.elapsed_time(e1) / 40 * 1e3:.1f} us")
ctx.set_option("small_waves", 8)
for L in (0, 4, 20):
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.FunnelPotential(D, 3.0), seed=1, bugCompat=False)
    st = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
    for stats in (None, st):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(23):
            if i == 3:
                e0.record()
            hmc.step(1 / KB, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        print(f"L={L} stats={'on' if stats is not None else 'off'}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us / iteration (device)")
    for adapt in (False,):
        hmc.run(3, 1 / KB)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hmc.run(50, 1 / KB, adapt=adapt)
        torch.cuda.synchronize()
        print(f"L={L} HMC.run: {(time.perf_counter() - t0) / 50 * 1e6:.1f} us / iteration (wall)")
# host-side cost of one step() call (enqueue only)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(200):
    hmc.step(1 / KB)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue cost of HMC.step: {(t1 - t0) / 200 * 1e6:.1f} us")
