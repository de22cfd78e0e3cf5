"""Config 1 (2-D isotropic Gaussian, P = 1024, L = 20, 1000 iterations) through the reference's own call,
HMC.getSamples, on a device ensemble: the whole loop as one launch (ehmc_hmc_run) vs one launch per iteration.
    python profiles/getsamples_probe.py
"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402

KB = 1.380649e-23
D, P, L, h, S = 2, 1024, 20, 0.05, 1000
for fused in (True, False):
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.HarmonicPotential(np.ones(D)), seed=1)
    if not fused:
        hmc._fused_loop_ok = lambda: False
    with contextlib.redirect_stdout(io.StringIO()):
        hmc.getSamples(S, 1 / KB, 1.0)  # warm-up of the same size (first-use cudaMalloc of the (D, P, S) arrays)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s, m = hmc.getSamples(S, 1 / KB, 1.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    x = s[:, :, 200:].double()
    print(f"fused={fused}: {dt * 1e3:.2f} ms for {S} iterations ({P * L * S / dt:.3e} particle-leapfrog-steps/s), "
          f"sample mean {x.mean().item():+.4f} var {x.var().item():.4f}")
