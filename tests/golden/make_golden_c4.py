"""Golden vector of BASELINE config 4 at FULL size (4096 bodies x 3-D per particle, ensemble of 1024, L = 10,
h = 0.01, Plummer eps = 0.05): one HMC iteration of 8 of the 1024 particles by the float64 oracle
(oracle/hmc_oracle.py: the reference's loop, src/integrator.py:105-120 + src/HMC.py:154-176, with the pairwise
potential whose -grad/m is pinned against the reference's getAccelNBody).  The oracle is O(B^2) NumPy -- 13 s per
particle -- so the test reads this fixture instead of recomputing it.

    python tests/golden/make_golden_c4.py        # writes tests/golden/c4_full_8.npz (0.4 MB)

Inputs are regenerated from the seed by the test (config4_inputs below is imported by it)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

B, P, L, H, EPS, SEED = 4096, 1024, 10, 0.01, 0.05, 404
SEL = np.array([0, 1, 127, 128, 511, 640, 1000, 1023])


def config4_inputs():
    rng = np.random.RandomState(SEED)
    q0 = rng.standard_normal((3 * B, P))
    z = rng.standard_normal((3 * B, P))
    u = rng.uniform(size=P)
    return q0, z, u


if __name__ == "__main__":
    from oracle import hmc_oracle as O

    q0, z, u = config4_inputs()
    m = np.ones(B) / B
    po = O.NBody(m, 1.0, EPS)
    qr, pr, accr, oh, nh = O.hmc_iter(q0[:, SEL], z[:, SEL], u[SEL], np.ones(len(SEL)), 1 / O.BOLTZMANN, H, L, po)
    np.savez_compressed(os.path.join(HERE, "c4_full_8.npz"), sel=SEL, q1=qr.astype(np.float32), accept=accr,
                        oldH=oh, newH=nh, q1_absmax=np.max(np.abs(qr)))
    print("accepted", accr, "oldH - newH", oh - nh)
