"""Times the UNMODIFIED reference (/root/reference/src, through oracle/ref_runner.py and the NumPy stand-in for the
nine jax.numpy names it uses) on BASELINE config 1 exactly: 2-D isotropic Gaussian, 1024 particles, L = 20,
h = 0.05 -- `iters` iterations of HMC.getSamples, extrapolated to the config's 1000.  TEST / MEASUREMENT
INFRASTRUCTURE: runs only in the build container (the GPU box has no /root/reference, and the reference is Python and
cannot travel); the result is committed as profiles/r02_ref_standin_c1.json and quoted by bench.py as
`cpu_baseline_ref` (kind "ref-standin"), marked with where it was measured.

    python oracle/time_reference_c1.py [iters]
"""
import json
import os
import platform
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_runner  # noqa: E402

KB = 1.380649e-23


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    D, P, L, h = 2, 1024, 20, 0.05
    k = np.ones(D)
    _, _, potential, _ = ref_runner.modules()
    pot = lambda q: potential.harmonicPotentialND(q, k)  # noqa: E731  (the reference's own potential function)
    grad = lambda q: k * q  # noqa: E731  (its analytic gradient; jax.grad is not available)
    ref_runner.run_get_samples(20221018, D, P, None, pot, grad, 2, 1 / KB, 1.0, 1.0, h)  # warm-up (imports)
    t0 = time.perf_counter()
    out = ref_runner.run_get_samples(20221018, D, P, None, pot, grad, iters, 1 / KB, 1.0, 1.0, h)
    dt = time.perf_counter() - t0
    assert out["numSteps"] == L
    res = {
        "config": "config1: 2-D isotropic Gaussian, P=1024, L=20, h=0.05",
        "kind": "ref-standin",
        "what": "unmodified /root/reference/src/{ensemble,integrator,potential,HMC}.py, HMC.getSamples, jax.numpy "
                "aliased to NumPy (oracle/jax_standin; JAX itself is not installable here), analytic gradient passed in",
        "iterations_timed": iters,
        "seconds": dt,
        "seconds_per_iteration": dt / iters,
        "extrapolated_seconds_for_1000_iterations": dt / iters * 1000,
        "value": P * L * iters / dt,
        "unit": "particle-leapfrog-steps/s",
        "cores": 1,
        "measured_on": f"build container CPU ({platform.processor() or platform.machine()}), python {platform.python_version()}, "
                       f"numpy {np.__version__}; NOT on the GPU box (the reference is Python and cannot travel)",
    }
    print(json.dumps(res, indent=1))
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                           "r02_ref_standin_c1.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
