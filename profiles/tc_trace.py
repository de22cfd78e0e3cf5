import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import physicsbasedbayesianinference_b200 as E
from physicsbasedbayesianinference_b200 import _lib
import bench
ctx = _lib.Context.get(0)
D, P, L, h = 100, 1 << 20, 50, 0.05
pot = E.GaussianPotential(precision=bench.make_precision(D))
ens = E.Ensemble(D, P, dtype=np.float32, device='cuda', seed=1)
ens.setPosition(1.0)
hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pot, seed=1, bugCompat=False)
for dbg in (0, 3):
    ctx.set_option('tc_debug', dbg)
    for _ in range(3): hmc.step(1 / 1.380649e-23)
    torch.cuda.synchronize()
    ctx.set_option('tc_prof', 1)
    hmc.step(1 / 1.380649e-23)
    torch.cuda.synchronize()
    print('dbg', dbg, file=sys.stderr)
    ctx.set_option('tc_prof_dump', 1)
    ctx.set_option('tc_prof', 0)
