"""CPU tests: the NumPy oracle (oracle/hmc_oracle.py) against the golden vectors
produced by the UNMODIFIED reference (tests/golden/make_golden.py) and against the
reference's own known answers KA1-KA3."""
import glob
import os

import numpy as np
import pytest

from oracle import hmc_oracle as O
from tests.conftest import GOLDEN, load_golden

KB = O.BOLTZMANN


def _pot(g, name):
    if "k" in g:
        return O.DiagGaussian(g["k"])
    if "prec" in g:
        return O.DenseGaussian(g["prec"], g["mean"])
    if "sigma_v" in g:
        return O.Funnel(int(g["D"]), float(g["sigma_v"]))
    if "body_mass" in g:
        return O.NBody(g["body_mass"], float(g["G"]), float(g["eps"]))
    if "X" in g:
        return O.Logistic(g["X"], g["y"], float(g["prior_scale"]))
    raise KeyError(name)


def test_known_answers():
    g = load_golden("known_answers")
    # KA1: src/tests/test_potential.py:20-25 -- exactly 33 in binary
    q = np.zeros((2, 10))
    q[:, 0] = (3.0, 4.0)
    assert O.DiagGaussian(np.array([2, 3])).energy(q)[0] == 33
    assert g["ka1"][0] == 33
    # KA2: fresh ensemble state
    assert np.all(g["ka2_q"] == 0) and np.all(g["ka2_p"] == 0) and g["ka2_m"] == 1.0 and g["ka2_w"] == 0.0
    assert g["ka2_index_error"] == 1
    # row F: int(finalTime/stepSize)
    cases = [(0.3, 0.1), (1.0, 0.05), (2.5, 0.05), (0.5, 0.05)]
    assert [O.num_steps(a, b) for a, b in cases] == list(g["num_steps"]) == [2, 20, 50, 10]


def test_get_accel_nbody_matches_reference():
    g = load_golden("known_answers")
    q, m = g["nbody_q"], g["nbody_m"]
    acc = np.stack([O.get_accel_nbody(q, m, i) for i in range(q.shape[1])], axis=1)
    assert np.array_equal(acc, g["nbody_acc"])  # same expression -> bit-exact
    # the NBody potential family: -grad_i U / m_i == getAccelNBody (eps = 0), SURVEY row N/N1
    nb = O.NBody(m, G=O.GRAV_CONST, eps=0.0)
    gr = nb.grad(q.reshape(-1)).reshape(3, -1)
    np.testing.assert_allclose(-gr / m, g["nbody_acc"], rtol=1e-12)
    # and U has the sign of samples/NBody/MiscFunctions.py:163-169, i.e. minus the reference's
    # (wrong-signed) nBodyPotential (row N1)
    np.testing.assert_allclose(nb.energy(q.reshape(-1)), -g["nbody_pot"], rtol=1e-13)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "integrate_*.npz"))))
def test_integrators_bit_exact(path):
    g = load_golden(os.path.basename(path)[:-4])
    pot = O.DiagGaussian(g["k"])
    L = O.num_steps(float(g["final_time"]), float(g["h"]))
    assert L == int(g["L"])
    f = O.stormer_verlet if "Stormer" in path else O.leapfrog
    q1, p1 = f(g["q0"], g["p0"], g["mass"], float(g["h"]), L, pot.grad)
    assert np.array_equal(q1, g["q1"]) and np.array_equal(p1, g["p1"])


def test_harmonic_analytic_KA3():
    """tests/test_integrator_harmonic.py:27-38, compared at numSteps*h:
    second-order convergence of the oracle leapfrog."""
    g = load_golden("integrate_Leapfrog_unitmass_h0.01")
    k, m = g["k"], g["mass"]
    errs = []
    for h in (1e-1, 1e-2, 1e-3):
        L = O.num_steps(float(g["final_time"]), h)
        t = L * h
        om = np.sqrt(np.outer(k, 1 / m))
        v0 = g["p0"] / m
        qa = g["q0"] * np.cos(om * t) + v0 / om * np.sin(om * t)
        q1, _ = O.leapfrog(g["q0"], g["p0"], m, h, L, O.DiagGaussian(k).grad)
        errs.append(np.max(np.abs(q1 - qa)) / np.max(np.abs(qa)))
    assert errs[0] / errs[1] == pytest.approx(100, rel=0.2)
    assert errs[1] / errs[2] == pytest.approx(100, rel=0.2)


@pytest.mark.parametrize("name", ["nbody_mode_Leapfrog"])
def test_nbody_reference_mode(name):
    g = load_golden(name)
    q, p = g["q0"], g["p0"]
    for c in range(g["q"].shape[0]):
        q, p = O.leapfrog_nbody_reference_mode(q, p, g["mass"], float(g["h"]), int(g["L"]))
        assert np.array_equal(q, g["q"][c]) and np.array_equal(p, g["p"][c])


HMC_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "hmc_*.npz")))


@pytest.mark.parametrize("name", HMC_CASES)
def test_hmc_get_samples(name):
    """Oracle getSamples loop fed the recorded z/u == reference samples and momenta
    (incl. the src/HMC.py:176 'p <- oldQ' quirk and the un-flipped stored momentum)."""
    g = load_golden(name)
    pot = _pot(g, name)
    method = "Stormer-Verlet" if "stormer" in name else "Leapfrog"
    q = g["z_init"] * float(g["q_std"])
    elementwise = isinstance(pot, (O.DiagGaussian,))
    n_rej = 0
    for it in range(int(g["S"])):
        q, p, acc, _, _ = O.hmc_iter(q, g["z"][it], g["u"][it], g["mass"], float(g["temperature"]),
                                     float(g["h"]), int(g["L"]), pot, method)
        n_rej += int((~acc).sum())
        if elementwise:
            assert np.array_equal(q, g["samples"][:, :, it])
            assert np.array_equal(p, g["momenta"][:, :, it])
        else:
            np.testing.assert_allclose(q, g["samples"][:, :, it], rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose(p, g["momenta"][:, :, it], rtol=1e-11, atol=1e-13)
    print(name, "rejections:", n_rej)


def test_get_samples_rng_stream_order():
    """seeding the global MT19937 and drawing in the reference's order
    reproduces the reference's samples without being fed z/u."""
    g = load_golden("hmc_iso2d")
    np.random.seed(20221018)
    s, m = O.get_samples(2, 64, g["mass"], O.DiagGaussian(g["k"]), int(g["S"]), float(g["temperature"]),
                         float(g["q_std"]), float(g["simul_time"]), float(g["h"]))
    assert np.array_equal(s, g["samples"]) and np.array_equal(m, g["momenta"])


@pytest.mark.parametrize("pot,D", [
    (O.DenseGaussian(np.array([[2.0, 0.3], [0.3, 1.0]]), np.array([1.0, -1.0])), 2),
    (O.Funnel(5, 3.0), 5),
    (O.NBody(np.array([0.2, 0.3, 0.5]), 1.0, 0.05), 9),
    (O.Logistic(np.random.RandomState(0).standard_normal((12, 4)),
                (np.random.RandomState(1).uniform(size=12) < 0.5).astype(float), 2.0), 4),
])
def test_closed_form_gradients_vs_finite_difference(pot, D):
    """The unpinned models: closed-form gradient == central difference of the energy
    (the reference's own fallback is scipy approx_fprime, src/potential.py:115-117)."""
    rng = np.random.RandomState(3)
    q = rng.standard_normal((D, 3))
    g = pot.grad(q)
    h = 1e-6
    for d in range(D):
        e = np.zeros((D, 1))
        e[d] = h
        fd = (pot.energy(q + e) - pot.energy(q - e)) / (2 * h)
        np.testing.assert_allclose(g[d], fd, rtol=2e-6, atol=1e-8)


def test_coin_toss_known_answer_KA6():
    """the reference's coin-toss sample (samples/NumpyroExamples/CoinToss/CoinToss.data.json:
    10 of 20 and 15 of 20 ones, reference biases p1 = 0.5, p2 = 0.75) -- the gradient of the log density
    vanishes at the reference biases (CoinTossExample.py:102-109; NotesOnParticleBasedHMC.pdf eq. 22)."""
    c1 = [1, 0] * 10
    c2 = [1] * 15 + [0] * 5
    pot = O.CoinToss([sum(c1), sum(c2)], [len(c1), len(c2)])
    assert np.array_equal(pot.grad(np.array([0.5, 0.75])), np.zeros(2))
    # and it is the minimum of U = -log density: 20 ln 2 + (-15 ln .75 - 5 ln .25)
    u_ref = 20 * np.log(2.0) - 15 * np.log(0.75) - 5 * np.log(0.25)
    assert abs(pot.energy(np.array([0.5, 0.75])) - u_ref) < 1e-12
    assert pot.energy(np.array([0.45, 0.75])) > u_ref and pot.energy(np.array([0.5, 0.8])) > u_ref
    # closed-form gradient == central difference inside (0, 1)
    q = np.array([[0.3, 0.6], [0.7, 0.2]])
    for d in range(2):
        e = np.zeros((2, 1))
        e[d] = 1e-6
        fd = (pot.energy(q + e) - pot.energy(q - e)) / 2e-6
        np.testing.assert_allclose(pot.grad(q)[d], fd, rtol=1e-6)


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10."""
    out = O.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    out = O.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(x) for x in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = O.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(x) for x in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_stream_moments():
    for dt in (np.float32, np.float64):
        z, u = O.philox_stream(1234, 7, np.arange(20000), 6, dt)
        assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
        assert abs(u.mean() - 0.5) < 0.01 and u.min() >= 0 and u.max() < 1


def test_philox_stream_version_2_uniform_rule():
    """Stream version 2 (include/ehmc.h, EHMC_RNG_STREAM_VERSION): float32 states with D mod 4 in {1, 2} take the
    Metropolis uniform from word z of the last normal block (block D // 4), every other case from word x (float64:
    x, y) of block 0xFFFFFFFF; the normals do not depend on the rule, and the uniform is independent of them."""
    seed, it = 0x0123456789ABCDEF, (3 << 32) | 17
    pid = np.arange(5000, dtype=np.uint64) + (1 << 33)
    lo, hi = pid & np.uint64(0xFFFFFFFF), pid >> np.uint64(32)
    k0, k1, itw = seed & 0xFFFFFFFF, ((seed >> 32) ^ (it >> 32)) & 0xFFFFFFFF, it & 0xFFFFFFFF
    assert O.PHILOX_STREAM_VERSION == 2
    for D in range(1, 13):
        z, u = O.philox_stream(seed, it, pid, D, np.float32)
        if D % 4 in (1, 2):
            _, _, w, _ = O.philox4x32_10(lo, hi, D // 4, itw, k0, k1)
        else:
            w, _, _, _ = O.philox4x32_10(lo, hi, O.PHILOX_UNIFORM_BLOCK, itw, k0, k1)
        assert np.array_equal(u, (w >> np.uint32(8)).astype(np.float64) * 2.0**-24)
        # the same dimensions of a wider state are the same normals (blocks are per 4 dimensions)
        z12, _ = O.philox_stream(seed, it, pid, 12, np.float32)
        assert np.array_equal(z, z12[:D])
        assert abs(np.corrcoef(u, z[D - 1])[0, 1]) < 0.05
    for D in (1, 2, 5):  # float64: all four words of a normal block are used, the uniform keeps its own block
        _, u = O.philox_stream(seed, it, pid, D, np.float64)
        x, y, _, _ = O.philox4x32_10(lo, hi, O.PHILOX_UNIFORM_BLOCK, itw, k0, k1)
        u53 = ((x >> np.uint32(6)).astype(np.uint64) << np.uint64(27)) | (y >> np.uint32(5)).astype(np.uint64)
        assert np.array_equal(u, u53.astype(np.float64) * 2.0**-53)


def test_ess_iid_and_correlated():
    rng = np.random.RandomState(5)
    x = rng.standard_normal((400, 64))
    e = O.ess_geyer(x)
    assert 0.7 * x.size < e < 1.3 * x.size
    y = np.zeros_like(x)
    for t in range(1, 400):
        y[t] = 0.9 * y[t - 1] + rng.standard_normal(64)
    assert O.ess_geyer(y) < 0.15 * y.size
