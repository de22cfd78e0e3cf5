"""Config 4 (4096 bodies x 3-D per particle, P = 1024, L = 10): bodies-per-thread variants of k_nbody.
    python profiles/nbody_probe.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402

KB = 1.380649e-23
B, P, L, h = 4096, 1024, 10, 0.01
ctx = E._lib.Context.get()
pot = E.NBodyPotential(np.ones(B) / B, G=1.0, eps=0.05).handle(32, ctx)
q = torch.randn(3 * B, P, device="cuda")
mass = torch.ones(P, device="cuda")
for ti in (8, 4):
    ctx.set_option("nbody_ti", ti)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(1 + 3):
        if i == 1:
            e0.record()
        args = E._lib.make_args(h, h * h, L, KB, 1 / KB, seed=1, iteration=i)
        E._lib.hmc_iter(ctx, pot, q, mass, args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"bodies/thread={ti}: {ms:.2f} ms / iteration, {P * (L + 1) * B * B / ms / 1e9:.0f} G interactions/s")
ctx.set_option("nbody_ti", 0)
