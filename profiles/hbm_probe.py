"""HBM fraction of the fused small-D trajectory kernel on its RNG-free entry points, P = 2^22 float32:
  * Leapfrog.integrate()  (reference row H): reads q, p, mass, writes q, p  -> (4 D + 1) * 4 bytes per particle
  * HMC.step with fed z, u (the parity harness): reads q, z, u, mass, writes q  -> (3 D + 2) * 4 bytes
  * HMC.step with in-kernel Philox (production): reads q, mass, writes q  -> (2 D + 1) * 4 bytes
Fractions are against MEASURED_PEAKS.json's copy bandwidth (fallback 6557 GB/s, the value measured on this pool).
    python profiles/hbm_probe.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physicsbasedbayesianinference_b200 as E  # noqa: E402

KB = 1.380649e-23
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6557.1
P, h = 1 << 22, 0.05


def timed(fn, n=20, warm=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warm + n):
        if i == warm:
            e0.record()
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def line(tag, D, L, t, nbytes):
    gbs = nbytes / t / 1e9
    print(f"{tag:<28} D={D:<3} L={L:<3} {t * 1e6:8.1f} us  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of copy peak "
          f"({nbytes / 1e6:.0f} MB)")


for name, mk in (("diag", lambda D: E.HarmonicPotential(np.linspace(1.0, 2.0, D))),
                 ("funnel", lambda D: E.FunnelPotential(D, 3.0))):
    for D in (4, 10, 16):
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
        ens.setPosition(1.0)
        ens.setMomentum(1 / KB)
        q0 = ens.q.clone()
        z = torch.randn(D, P, device="cuda")
        u = torch.rand(P, device="cuda")
        for L in (1, 4, 8, 20):
            pot = mk(D)
            lf = E.Leapfrog(ens, h, L * h + 1e-9, pot)
            ens.q.copy_(q0)
            line(f"{name} Leapfrog.integrate", D, L, timed(lf.integrate), (4 * D + 1) * 4 * P)
            hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pot, seed=1, bugCompat=False)
            ens.q.copy_(q0)
            line(f"{name} HMC.step fed z,u", D, L, timed(lambda: hmc.step(1 / KB, z=z, u=u)), (3 * D + 2) * 4 * P)
            ens.q.copy_(q0)
            line(f"{name} HMC.step Philox", D, L, timed(lambda: hmc.step(1 / KB)), (2 * D + 1) * 4 * P)
