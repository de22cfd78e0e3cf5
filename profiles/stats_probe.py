"""Kernel-level cost of the ensemble statistics and of the fused ensemble run (run under
`ncu --metrics gpu__time_duration.sum`): k_small with / without statistics and k_small_ens for a fixed number of
iterations at adaptLag 1 and 2, P = 2^19 and 2^22, funnel 10-D."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import physicsbasedbayesianinference_b200 as E

KB = 1.380649e-23
D, h = 10, 0.3
L = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
for logP in (19, 22):
    P = 1 << logP
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
    ens.setPosition(1.0)
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.FunnelPotential(D, 3.0), seed=1, bugCompat=False)
    hmc.run(50, 1 / KB, adapt=True, keepNumSteps=True)
    st = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
    for _ in range(3):
        hmc.step(1 / KB)
    for _ in range(3):
        hmc.step(1 / KB, stats=st)
    for lag in (1, 2):
        hmc.run(N, 1 / KB, adapt=False, keepNumSteps=True, adaptLag=lag)
    torch.cuda.synchronize()
