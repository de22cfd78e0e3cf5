"""Times the UNMODIFIED reference (/root/reference/src through oracle/ref_runner.py and the NumPy stand-in) on BASELINE
configs 2 and 5 at REDUCED ensemble size (CPU baseline item 1): the reference is a serial Python
loop over particles and leapfrog steps, so its rate per particle-leapfrog-step does not depend on P.
  config 2: 100-D dense Gaussian (the bench's precision matrix), P = 64,  L = 50, h = 0.05
  config 5: Neal's funnel 10-D,                                  P = 256, L = 20, h = 0.05
The model gradients are the closed forms of oracle/hmc_oracle.py evaluated per particle (the reference calls
gradient(q[:, i]); jax.grad is not available).  TEST / MEASUREMENT INFRASTRUCTURE: runs only in the build container;
the result is committed as profiles/r02_ref_standin_reduced.json and quoted by bench.py inside `cpu_baseline_ref`.

    python oracle/time_reference_reduced.py
"""
import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_runner  # noqa: E402

KB = 1.380649e-23
SEED = 20221018


def time_it(name, D, P, L, h, pot, grad, iters):
    ref_runner.run_get_samples(SEED, D, min(P, 8), None, pot, grad, 1, 1 / KB, 1.0, L * h + 1e-9, h)  # warm-up
    t0 = time.perf_counter()
    out = ref_runner.run_get_samples(SEED, D, P, None, pot, grad, iters, 1 / KB, 1.0, L * h + 1e-9, h)
    dt = time.perf_counter() - t0
    assert out["numSteps"] == L and np.isfinite(out["samples"]).all()
    return {"config": name, "D": D, "P": P, "L": L, "h": h, "iterations_timed": iters, "seconds": dt,
            "value": P * L * iters / dt, "unit": "particle-leapfrog-steps/s", "cores": 1}


def main():
    rng = np.random.RandomState(SEED)
    D = 100
    A = rng.standard_normal((D, D))
    lam = A @ A.T / D + np.eye(D)  # == bench.make_precision(100)
    c2 = time_it("config2 at reduced P: 100-D dense-precision Gaussian", D, 64, 50, 0.05,
                 lambda q: 0.5 * q @ lam @ q, lambda q: lam @ q, 20)

    def funnel_u(q):
        return q[0] ** 2 / 18.0 + 0.5 * np.exp(-q[0]) * np.sum(q[1:] ** 2) + 4.5 * q[0]

    def funnel_g(q):
        ev = np.exp(-q[0])
        g = ev * q
        g[0] = q[0] / 9.0 - 0.5 * ev * np.sum(q[1:] ** 2) + 4.5
        return g

    c5 = time_it("config5 at reduced P: Neal's funnel 10-D", 10, 256, 20, 0.05, funnel_u, funnel_g, 20)
    res = {"kind": "ref-standin", "what": "unmodified /root/reference/src HMC.getSamples under the jax.numpy -> NumPy "
           "stand-in, closed-form gradients passed in; serial Python loop over particles, so the rate per "
           "particle-leapfrog-step is independent of P",
           "c2": c2, "c5": c5,
           "measured_on": f"build container CPU ({platform.processor() or platform.machine()}), python "
                          f"{platform.python_version()}, numpy {np.__version__}; NOT on the GPU box"}
    print(json.dumps(res, indent=1))
    with open(os.path.join(ROOT, "profiles", "r02_ref_standin_reduced.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
