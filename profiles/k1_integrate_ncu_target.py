"""ncu target: Leapfrog.integrate (reference row H) on the small-D kernel, funnel D = 10, P = 2^22, L = 4, float32.
    ncu --set full --clock-control none --import-source on -k regex:k_small -s 3 -c 1 python profiles/k1_integrate_ncu_target.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402

D, P, h, L = 10, 1 << 22, 0.05, 4
ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
ens.setPosition(1.0)
ens.setMomentum(1 / 1.380649e-23)
lf = E.Leapfrog(ens, h, L * h + 1e-9, E.FunnelPotential(D, 3.0))
for i in range(6):
    lf.integrate()
torch.cuda.synchronize()
print("ok")
