"""clock64 phase trace of one steady-state tile of k_dense_tc3 (CTA 0, group 0, thread 0) at config 2.
Stamps: 0 tile start | 1 position loads issued | 2 momenta drawn | 3 positions parked, momenta in
registers | 4 operands split | after evaluations 0, 1, L-1, L | Metropolis decided | write-back issued.
    python profiles/tc_trace.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402
import bench  # noqa: E402

ctx = E._lib.Context.get(0)
D, P, h = 100, 1 << 20, 0.05
pot = E.GaussianPotential(precision=bench.make_precision(D))
ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
ens.setPosition(1.0)
for L in (50, 0):
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pot, seed=1, bugCompat=False)
    for dbg in (0, 3):
        ctx.set_option("tc_debug", dbg)
        for _ in range(3):
            hmc.step(1 / 1.380649e-23)
        torch.cuda.synchronize()
        ctx.set_option("tc_prof", 1)
        hmc.step(1 / 1.380649e-23)
        torch.cuda.synchronize()
        print("L", L, "dbg", dbg, file=sys.stderr)
        ctx.set_option("tc_prof_dump", 1)
        ctx.set_option("tc_prof", 0)
ctx.set_option("tc_debug", 0)
