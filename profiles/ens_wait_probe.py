"""Where the fused ensemble run (k_small_ens) waits at the 8-GPU shard size of config 5 (2^19 particles on ONE GPU):
total time the compute warps spin for their group's mark / for the published step size, master phase stamps, for
adaptLag 1 and 2."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import physicsbasedbayesianinference_b200 as E

KB = 1.380649e-23
D, L, h = 10, int(sys.argv[2]) if len(sys.argv) > 2 else 20, 0.05
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = E._lib.Context.get()
for logP in (19, 22):
    for lag in (2, 1):
        P = 1 << logP
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
        ens.setPosition(1.0)
        hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.FunnelPotential(D, 3.0), seed=1, bugCompat=False)
        hmc.run(50, 1 / KB, adapt=True, keepNumSteps=True)
        torch.cuda.synchronize()
        ctx.set_option("ens_debug", 24)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hmc.run(n, 1 / KB, adapt=False, keepNumSteps=True, adaptLag=lag)
        e1.record()
        torch.cuda.synchronize()
        print(f"P=2^{logP} L={L} lag={lag}: {1e3 * e0.elapsed_time(e1) / n:.2f} us/iteration over {n} iterations "
              f"({1e-3 * e0.elapsed_time(e1):.4f} s)", file=sys.stderr, flush=True)
        ctx.set_option("ens_debug_dump", 12)
        ctx.set_option("ens_debug", 0)
