"""Config 3 (X 100k x 256, P = 65536): one gradient launch of k_logistic_tc with and without the energy.
    python profiles/logistic_probe.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicsbasedbayesianinference_b200 as E  # noqa: E402
import bench  # noqa: E402

N, D, P = 100000, 256, 65536
X, y = bench.make_logistic_data(D, N)
pot = E.LogisticPotential(X, y, 1.0, precision="bf16")
ctx = E._lib.Context.get()
hd = pot.handle(32, ctx)
th = torch.randn(D, P, device="cuda") * 0.1
g = torch.empty_like(th)
e = torch.empty(P, device="cuda")


def t(fn, n=5):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


flop = 4.0 * N * D * P
for name, fn in (("grad only", lambda: E._lib.potential_eval(ctx, hd, th, None, g)),
                 ("grad + energy", lambda: E._lib.potential_eval(ctx, hd, th, e, g))):
    ms = t(fn)
    print(f"{name}: {ms:.3f} ms  {flop / ms / 1e9:.0f} TFLOP/s")
