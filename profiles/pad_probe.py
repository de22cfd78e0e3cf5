"""Small-D kernel at dimensions padded up to the instantiated width (D = 9 -> 10, 3 -> 4) next to the exact widths:
funnel, P = 2^22, float32, Leapfrog.integrate and HMC.step with in-kernel Philox.  Run from the repository root:
    python profiles/pad_probe.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import physicsbasedbayesianinference_b200 as E
KB = 1.380649e-23
P, h = 1 << 22, 0.05
for D in (9, 10, 3, 4):
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
    ens.setPosition(1.0); ens.setMomentum(1 / KB)
    for L in (4, 20):
        pot = E.FunnelPotential(D, 3.0)
        lf = E.Leapfrog(ens, h, L * h + 1e-9, pot)
        hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pot, seed=1, bugCompat=False)
        for tag, fn in (("integrate", lf.integrate), ("HMC.step Philox", lambda: hmc.step(1 / KB))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(23):
                if i == 3: e0.record()
                fn()
            e1.record(); torch.cuda.synchronize()
            print(f"funnel D={D} L={L} {tag}: {e0.elapsed_time(e1)/20*1e3:.1f} us")
