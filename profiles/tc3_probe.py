"""Where does an iteration of k_dense_tc3 go?  Times config 2 (100-D, P = 2^20) for several
trajectory lengths with the kernel's debug knobs (1 = no MMAs, 2 = no epilogue arithmetic) and
with fed momenta (no Philox / Box-Muller in the prologue).  Run on the GPU box:
    python profiles/tc3_probe.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402

KB = 1.380649e-23
D, P, h = 100, 1 << 20, 0.05
rng = np.random.RandomState(20221018)
A = rng.standard_normal((D, D))
prec = A @ A.T / D + np.eye(D)
ctx = E._lib.Context.get()
pot = E.GaussianPotential(precision=prec).handle(32, ctx)
q = torch.randn(D, P, device="cuda")
mass = torch.ones(P, device="cuda")
z = torch.randn(D, P, device="cuda")
u = torch.rand(P, device="cuda")


def time_it(L, dbg, fed, path=4, iters=10):
    ctx.set_option("dense_path", path)
    ctx.set_option("tc_debug", dbg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3 + iters):
        if i == 3:
            e0.record()
        args = E._lib.make_args(h, h * h, L, KB, 1 / KB, seed=1, iteration=i)
        E._lib.hmc_iter(ctx, pot, q, mass, args, z=z if fed else None, u=u if fed else None)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print("path L dbg fed ms")
for path in (4, 3):
    for L in (0, 10, 50):
        for dbg in (0, 1, 2, 3):
            for fed in (False, True):
                if (path == 3 and (dbg >= 3 or fed)) or (fed and dbg):
                    continue
                q.normal_()
                print(path, L, dbg, int(fed), "%.3f" % time_it(L, dbg, fed, path), flush=True)
ctx.set_option("tc_debug", 0)
ctx.set_option("dense_path", 0)
