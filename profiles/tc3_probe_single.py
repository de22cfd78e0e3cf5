"""Single-group latencies of k_dense_tc3: P = 148 * 128 particles gives every SM exactly one tile,
so nothing overlaps and (T(L=50) - T(L=0)) / 50 is the serial time of one evaluation:
dbg 0 = MMAs + epilogue + sync, 1 = epilogue + sync, 2 = MMAs + sync, 3 = sync only.
    python profiles/tc3_probe_single.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402

KB = 1.380649e-23
D, h = 100, 0.05
rng = np.random.RandomState(20221018)
A = rng.standard_normal((D, D))
prec = A @ A.T / D + np.eye(D)
ctx = E._lib.Context.get()
pot = E.GaussianPotential(precision=prec).handle(32, ctx)


def time_it(P, L, dbg, iters=20):
    q = torch.randn(D, P, device="cuda")
    mass = torch.ones(P, device="cuda")
    ctx.set_option("dense_path", 4)
    ctx.set_option("tc_debug", dbg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3 + iters):
        if i == 3:
            e0.record()
        args = E._lib.make_args(h, h * h, L, KB, 1 / KB, seed=1, iteration=i)
        E._lib.hmc_iter(ctx, pot, q, mass, args)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


for P in (148 * 128, 148 * 256):
    for dbg in (0, 1, 2, 3):
        t0, t1 = time_it(P, 0, dbg), time_it(P, 200, dbg)
        print(f"P={P} dbg={dbg}: L=0 {t0:.1f} us, L=200 {t1:.1f} us, per evaluation {(t1 - t0) / 200 * 1e3:.0f} ns")
ctx.set_option("tc_debug", 0)
ctx.set_option("dense_path", 0)
