"""Debug aid: HMC.step(stats=...) (k_small with statistics + k_stats_finalize, two launches) at 2^22 particles was
seen at 0.13-0.27 ms per iteration in some probe runs and at 0.7-3 ms in others.  Segments of 50 iterations: CPU
enqueue time and GPU time of each, with and without a synchronize between segments."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import physicsbasedbayesianinference_b200 as E

KB = 1.380649e-23
D, L, h, P = 10, int(sys.argv[1]) if len(sys.argv) > 1 else 20, 0.3, 1 << 22
ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
ens.setPosition(1.0)
hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.FunnelPotential(D, 3.0), seed=1, bugCompat=False)
hmc.run(50, 1 / KB, adapt=True, keepNumSteps=True)
st = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
for mode in ("bare", "stats", "stats", "bare", "stats"):
    torch.cuda.synchronize()
    rows = []
    for seg in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(50):
            if mode == "bare":
                hmc.step(1 / KB)
            else:
                hmc.step(1 / KB, stats=st)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        rows.append((1e6 * (t1 - t0) / 50, 1e3 * e0.elapsed_time(e1) / 50))
    print(mode, " ".join(f"[cpu {c:.0f} us, gpu {g:.0f} us]" for c, g in rows), flush=True)
print("acceptance", float(st[0]) / P, "finite", bool(torch.isfinite(ens.q).all()))
