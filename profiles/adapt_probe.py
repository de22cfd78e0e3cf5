"""Config 5 (funnel, D = 10, P = 2^22, L = 20): per-iteration wall time of the adaptive loop --
host-side adapter (stale statistics through a side stream) vs device-resident control blocks, eager and
as a CUDA graph.
    python profiles/adapt_probe.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402

KB = 1.380649e-23
D, h, L = 10, 0.05, 20
for P in (1 << 22, 1 << 19):
    for mode in ("host", "device", "device+graph"):
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=1)
        ens.setPosition(1.0)
        hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.FunnelPotential(D, 3.0), seed=1, bugCompat=False)
        kw = dict(adapt=True, keepNumSteps=True, deviceAdapt=mode != "host", graph=mode.endswith("graph"))
        hmc.run(12, 1 / KB, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hmc.run(102, 1 / KB, **kw)
        torch.cuda.synchronize()
        print(f"P=2^{P.bit_length() - 1} {mode:13s}: {(time.perf_counter() - t0) / 102 * 1e6:7.1f} us / iteration, step size {hmc.stepSize:.6f}")
