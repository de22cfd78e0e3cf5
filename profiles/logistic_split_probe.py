"""Accuracy of the fp16-split tensor-core logistic gradient (k_logistic_tcs) against the float64 oracle as the
number of data rows N (= length of the fp32 accumulation chain of GEMM2 in tensor memory) grows."""
import sys

import numpy as np

sys.path.insert(0, ".")
import physicsbasedbayesianinference_b200 as E
from oracle import hmc_oracle as O

D, P = 256, 256
rng = np.random.RandomState(1)
th_true = rng.standard_normal(D)
for N in (1000, 8000, 32000, 100000):
    X = rng.standard_normal((N, D)) / np.sqrt(D)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-X @ th_true))).astype(np.float64)
    for name, th in (("prior-draw", rng.standard_normal((D, P))),
                     ("near-posterior", th_true[:, None] + 0.1 * rng.standard_normal((D, P)))):
        th32 = th.astype(np.float32)
        po = O.Logistic(X, y, 1.0)
        g_ref, u_ref = po.grad(th32.astype(np.float64)), po.energy(th32.astype(np.float64))
        for prec in ("fp16x3", "fp32"):
            pe = E.LogisticPotential(X, y, 1.0, precision=prec)
            g, u = pe.gradient(th32), pe(th32)
            eg = np.max(np.abs(g - g_ref), axis=0) / np.max(np.abs(g_ref), axis=0)
            # signed relative error along the gradient direction: a rounding bias shows up here
            bias = np.sum((g - g_ref) * g_ref, axis=0) / np.sum(g_ref * g_ref, axis=0)
            eu = np.abs(u - u_ref)
            print(f"N={N:6d} {name:14s} {prec:7s} grad rel err median {np.median(eg):.2e} max {eg.max():.2e} "
                  f"bias {np.median(bias):+.2e}   |U| {np.median(np.abs(u_ref)):.3e} abs err U median "
                  f"{np.median(eu):.2e} max {eu.max():.2e}", flush=True)
