"""ncu target: the fused adaptive ensemble run (k_small_ens) at a given shard size on one GPU.
    ncu --set full -k regex:k_small_ens -s 1 -c 1 ... python profiles/ens_ncu_target.py <log2 P> <L> <iterations>"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import physicsbasedbayesianinference_b200 as E

KB = 1.380649e-23
logP, L, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
D, h = 10, 0.05
ens = E.Ensemble(D, 1 << logP, dtype=np.float32, device="cuda", seed=1)
ens.setPosition(1.0)
hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.FunnelPotential(D, 3.0), seed=1, bugCompat=False)
hmc.run(50, 1 / KB, adapt=True, keepNumSteps=True)   # launch 0: warm-up with adaptation
hmc.run(n, 1 / KB, adapt=False, keepNumSteps=True, adaptLag=2)  # launch 1: the captured one
torch.cuda.synchronize()
print("ok", hmc.stepSize)
