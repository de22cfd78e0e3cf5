"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C-ABI
(ctypes -> _ehmc.so), against
  * the golden vectors produced by the UNMODIFIED reference (tests/golden/*.npz),
  * the NumPy oracle on the same seeded inputs,
  * size-independent properties at BASELINE.json's full sizes.

Tolerances (north_star): trajectories within 1e-5 relative in float32 and 1e-12 in
float64 ("relative" = max |a-b| / max |b| over the compared array); accept/reject
decisions identical away from threshold ties (|u - min(1, ratio)| < TIE).
"""
import glob
import os

import numpy as np
import pytest

from oracle import hmc_oracle as O
from tests.conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

RTOL = {np.float32: 1e-5, np.float64: 1e-12}
TIE = {np.float32: 2e-3, np.float64: 1e-9}
KB = O.BOLTZMANN


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def E():
    import physicsbasedbayesianinference_b200 as pkg

    pkg._lib.Context.get()  # fails loudly when the extension or the GPU is missing
    return pkg


def descriptor(E, g):
    """(engine descriptor, oracle potential) for a golden case."""
    if "k" in g:
        return E.HarmonicPotential(g["k"]), O.DiagGaussian(g["k"])
    if "prec" in g:
        return E.GaussianPotential(precision=g["prec"], mean=g["mean"]), O.DenseGaussian(g["prec"], g["mean"])
    if "sigma_v" in g:
        return E.FunnelPotential(int(g["D"]), float(g["sigma_v"])), O.Funnel(int(g["D"]), float(g["sigma_v"]))
    if "body_mass" in g:
        return (E.potential.NBodyPotential(g["body_mass"], G=float(g["G"]), eps=float(g["eps"])),
                O.NBody(g["body_mass"], float(g["G"]), float(g["eps"])))
    if "X" in g:
        return (E.potential.LogisticPotential(g["X"], g["y"], float(g["prior_scale"])),
                O.Logistic(g["X"], g["y"], float(g["prior_scale"])))
    raise KeyError


# ---------------------------------------------------------------------------
# known answers of the reference's own tests
# ---------------------------------------------------------------------------
def test_KA1_harmonic_potential_is_33(E):
    """src/tests/test_potential.py:14-25 through the CUDA evaluation kernel."""
    ens = E.Ensemble(2, 10)
    ens.q[:, 0] = np.array([3.0, 4.0])
    pot = E.harmonicPotentialND(ens.q, np.array([2, 3]))
    assert pot[0] == 33
    assert np.all(pot[1:] == 0)
    assert E.harmonicPotentialND(np.array([3.0, 4.0]), np.array([2.0, 3.0])) == 33


def test_potential_eval_matches_oracle(E):
    rng = np.random.RandomState(1)
    cases = [
        (E.HarmonicPotential([2.0, 3.0, 4.0]), O.DiagGaussian([2.0, 3.0, 4.0]), 3),
        (E.FunnelPotential(10, 3.0), O.Funnel(10, 3.0), 10),
    ]
    A = rng.standard_normal((20, 20))
    prec = A @ A.T / 20 + np.eye(20)
    mu = rng.standard_normal(20)
    cases.append((E.GaussianPotential(precision=prec, mean=mu), O.DenseGaussian(prec, mu), 20))
    prec2 = np.linalg.inv(np.array([[4.0, -3.0], [-3.0, 4.0]]))
    cases.append((E.GaussianPotential(precision=prec2, mean=[5.0, 5.0]), O.DenseGaussian(prec2, [5.0, 5.0]), 2))
    for pe, po, D in cases:
        q = rng.standard_normal((D, 37))
        for dt in (np.float64, np.float32):
            qq = q.astype(dt)
            assert rel_err(pe(qq), po.energy(q)) < 10 * RTOL[dt]
            assert rel_err(pe.gradient(qq), po.grad(q)) < 10 * RTOL[dt]
        # (D,) input -> scalar / (D,) like the reference's per-particle callables
        assert np.shape(pe(q[:, 0])) == ()
        assert pe.gradient(q[:, 0]).shape == (D,)


# ---------------------------------------------------------------------------
# integrators vs the reference's golden vectors
# ---------------------------------------------------------------------------
INTEG = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "integrate_*.npz")))


@pytest.mark.parametrize("name", INTEG)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("backing", ["host", "device"])
def test_integrate_golden(E, name, dt, backing):
    import torch

    g = load_golden(name)
    D, P = g["q0"].shape
    cls = E.StormerVerlet if "Stormer" in name else E.Leapfrog
    if backing == "host":
        ens = E.Ensemble(D, P, dtype=dt)
        ens.mass = g["mass"].astype(dt)
        ens.q[:] = g["q0"]
        ens.p[:] = g["p0"]
    else:
        ens = E.Ensemble(D, P, dtype=dt, device="cuda")
        tdt = torch.float64 if dt == np.float64 else torch.float32
        ens.mass = torch.tensor(g["mass"], dtype=tdt, device="cuda")
        ens.q.copy_(torch.tensor(g["q0"], dtype=tdt))
        ens.p.copy_(torch.tensor(g["p0"], dtype=tdt))
    pot = E.HarmonicPotential(g["k"])
    integ = cls(ens, float(g["h"]), float(g["final_time"]), pot.gradient)
    assert integ.numSteps == int(g["L"])
    q, p = integ.integrate()
    assert q is ens.q and p is ens.p  # in place, same objects
    qn = q if backing == "host" else q.cpu().numpy()
    pn = p if backing == "host" else p.cpu().numpy()
    # float32 and long trajectories (h = 0.01: 444 steps): rounding accumulates ~sqrt(L) for the
    # velocity form and ~L for the two-step position Verlet recurrence (q = 2q - qPast + a h^2)
    grow = int(g["L"]) / 50.0
    tol = RTOL[dt] * (max(1.0, grow if "Stormer" in name else np.sqrt(grow)) if dt == np.float32 else 1.0)
    assert rel_err(qn, g["q1"]) < tol
    assert rel_err(pn, g["p1"]) < tol
    if dt == np.float64:
        # element-wise family in float64: the kernel evaluates the reference's expressions in
        # the reference's order without FMA contraction -> bit-identical
        assert np.array_equal(qn, g["q1"]) and np.array_equal(pn, g["p1"])


def test_harmonic_analytic_KA3(E):
    """tests/test_integrator_harmonic.py:27-38 at numSteps*h: O(h^2) convergence on the GPU."""
    g = load_golden("integrate_Leapfrog_unitmass_h0.01")
    k, m = g["k"], g["mass"]
    errs = []
    for h in (1e-1, 1e-2, 1e-3):
        ens = E.Ensemble(2, 16)
        ens.q[:] = g["q0"]
        ens.p[:] = g["p0"]
        integ = E.Leapfrog(ens, h, float(g["final_time"]), E.HarmonicPotential(k))
        t = integ.numSteps * h
        om = np.sqrt(np.outer(k, 1 / m))
        qa = g["q0"] * np.cos(om * t) + g["p0"] / m / om * np.sin(om * t)
        q, _ = integ.integrate()
        errs.append(rel_err(q, qa))
    assert errs[0] / errs[1] == pytest.approx(100, rel=0.2)
    assert errs[1] / errs[2] == pytest.approx(100, rel=0.2)


@pytest.mark.parametrize("method", ["Leapfrog", "StormerVerlet"])
def test_nbody_reference_mode(E, method):
    """Integrator(..., gradient=None): Sun/Earth/Moon of tests/test_integrator_solar_system.py."""
    g = load_golden("nbody_mode_" + method)
    ens = E.Ensemble(3, 3)
    ens.mass = g["mass"].copy()
    ens.q[:] = g["q0"]
    ens.p[:] = g["p0"]
    cls = E.Leapfrog if method == "Leapfrog" else E.StormerVerlet
    integ = cls(ens, float(g["h"]), float(g["final_time"]), None)
    for c in range(g["q"].shape[0]):
        q, p = integ.integrate()
        assert rel_err(q, g["q"][c]) < 1e-12
        assert rel_err(p, g["p"][c]) < 1e-10


# ---------------------------------------------------------------------------
# HMC iterations vs the reference's golden chains (fed the reference's own z / u)
# ---------------------------------------------------------------------------
HMC_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "hmc_*.npz")))


def _supported(E, g):
    try:
        return descriptor(E, g)
    except AttributeError:
        return None


@pytest.mark.parametrize("name", HMC_CASES)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_hmc_golden_single_iterations(E, name, dt):
    """Every iteration started from the REFERENCE's state: q, stored momentum and accept
    decision of one fused-kernel iteration vs the reference (host path of the C-ABI)."""
    g = load_golden(name)
    d = _supported(E, g)
    if d is None:
        pytest.skip("family not built yet")
    pe, po = d
    D, P, S, L = int(g["D"]), int(g["P"]), int(g["S"]), int(g["L"])
    method = "Stormer-Verlet" if "stormer" in name else "Leapfrog"
    ens = E.Ensemble(D, P, dtype=dt)
    ens.mass = g["mass"].astype(dt)
    hmc = E.HMC(ens, float(g["simul_time"]), float(g["h"]), None, potential=pe, method=method)
    assert hmc.integrator.numSteps == L
    q_prev = g["z_init"] * float(g["q_std"])
    n_tie = 0
    for it in range(S):
        z, u = g["z"][it], g["u"][it]
        _, _, acc_ref, oldH, newH = O.hmc_iter(q_prev, z, u, g["mass"], float(g["temperature"]), float(g["h"]), L, po,
                                               method)
        hmc.integrator.q = np.ascontiguousarray(q_prev, dtype=dt)
        p_out = np.empty((D, P), dtype=dt)
        acc = np.empty(P, dtype=np.uint8)
        hmc.step(float(g["temperature"]), p_out=p_out, accept=acc, z=z.astype(dt), u=u.astype(dt))
        with np.errstate(over="ignore"):
            margin = np.abs(u - np.minimum(1, np.exp(oldH - newH)))
        clear = margin > TIE[dt]
        assert np.array_equal(acc.astype(bool)[clear], acc_ref[clear]), "accept/reject differs away from a tie"
        same = acc.astype(bool) == acc_ref
        n_tie += int((~same).sum())
        assert rel_err(hmc.integrator.q[:, same], g["samples"][:, same, it]) < RTOL[dt]
        assert rel_err(p_out[:, same], g["momenta"][:, same, it]) < RTOL[dt]
        q_prev = g["samples"][:, :, it]
    assert n_tie <= 1


@pytest.mark.parametrize("name", HMC_CASES)
def test_hmc_golden_chain_float64(E, name):
    """Whole chain in float64, chained on the GPU's own state, must track the reference."""
    g = load_golden(name)
    d = _supported(E, g)
    if d is None:
        pytest.skip("family not built yet")
    pe, _ = d
    D, P, S = int(g["D"]), int(g["P"]), int(g["S"])
    method = "Stormer-Verlet" if "stormer" in name else "Leapfrog"
    ens = E.Ensemble(D, P)
    ens.mass = g["mass"].copy()
    hmc = E.HMC(ens, float(g["simul_time"]), float(g["h"]), None, potential=pe, method=method)
    hmc.integrator.q = np.ascontiguousarray(g["z_init"] * float(g["q_std"]))
    for it in range(S):
        p_out = np.empty((D, P))
        hmc.step(float(g["temperature"]), p_out=p_out, z=g["z"][it].copy(), u=g["u"][it].copy())
        assert rel_err(hmc.integrator.q, g["samples"][:, :, it]) < 1e-11
        assert rel_err(p_out, g["momenta"][:, :, it]) < 1e-11


def test_getSamples_drop_in_numpy_stream(E, capsys):
    """np.random.seed(s); HMC(...).getSamples(...) reproduces the reference's chain:
    same RNG stream, same print-out, same (D,P,S) arrays (tests/golden/hmc_iso2d)."""
    g = load_golden("hmc_iso2d")
    np.random.seed(20221018)
    ens = E.Ensemble(2, 64)
    hmc = E.HMC(ens, float(g["simul_time"]), float(g["h"]), None, potential=E.HarmonicPotential(g["k"]))
    samples, momenta = hmc.getSamples(int(g["S"]), float(g["temperature"]), float(g["q_std"]))
    out = capsys.readouterr().out
    assert "HMC iteration  1" in out and "time step:  0.05" in out
    assert samples.shape == (2, 64, 12) and samples.dtype == np.float64
    assert rel_err(samples, g["samples"]) < 1e-12
    assert rel_err(momenta, g["momenta"]) < 1e-12


@pytest.mark.parametrize("case", ["iso2", "funnel10", "dense8", "coin2"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("bug", [True, False])
def test_getSamples_fused_loop_equals_iteration_by_iteration(E, case, dt, bug, capsys):
    """getSamples on a device ensemble with Philox draws runs the whole loop of src/HMC.py:150-179 as ONE
    launch (ehmc_hmc_run); it must reproduce the per-iteration path (one ehmc_hmc_iter per iteration,
    samples copied out in between) bit for bit, including the stored momenta of rejected particles."""
    import torch

    rng = np.random.RandomState(31)
    P, S = 777, 23
    if case == "iso2":
        D, pot, h, T = 2, E.HarmonicPotential([1.0, 2.5]), 0.3, 0.9
    elif case == "funnel10":
        D, pot, h, T = 10, E.FunnelPotential(10, 3.0), 0.2, 1.0
    elif case == "coin2":
        D, pot, h, T = 2, E.CoinTossPotential([10, 15], [20, 20]), 0.02, 0.2
    else:
        D = 8
        A = rng.standard_normal((D, D))
        pot, h, T = E.GaussianPotential(precision=A @ A.T / D + np.eye(D), mean=rng.standard_normal(D)), 0.25, 1.0
    q0 = rng.uniform(0.3, 0.7, (D, P)) if case == "coin2" else rng.standard_normal((D, P))
    mass = rng.uniform(0.5, 2.0, P)
    tdt = torch.float32 if dt == np.float32 else torch.float64

    def fresh(method):
        ens = E.Ensemble(D, P, dtype=dt, device="cuda", seed=9)
        ens.mass = torch.tensor(mass, dtype=tdt, device="cuda")
        hmc = E.HMC(ens, T, h, None, potential=pot, seed=9, bugCompat=bug, method=method)
        ens.setPosition = lambda qStd: ens.q.copy_(torch.tensor(q0, dtype=tdt))  # fixed start (copy_ returns ens.q)
        return ens, hmc

    for method in ("Leapfrog", "Stormer-Verlet"):
        ens_a, hmc_a = fresh(method)
        sa, ma = hmc_a.getSamples(S, 1 / KB, 1.0)  # fused
        ens_b, hmc_b = fresh(method)
        hmc_b._fused_loop_ok = lambda: False
        sb, mb = hmc_b.getSamples(S, 1 / KB, 1.0)  # one launch per iteration
        assert hmc_a.iteration == hmc_b.iteration == S
        assert torch.equal(sa, sb), (case, method)
        assert torch.equal(ma, mb), (case, method)
        assert torch.equal(ens_a.q, ens_b.q) and torch.equal(ens_a.p, ens_b.p)
        moved = (sa[:, :, -1] != torch.tensor(q0, dtype=tdt, device="cuda")).any(0).float().mean().item()
        assert moved > 0.5  # the chain actually accepts proposals
    capsys.readouterr()


def test_bugcompat_flag(E):
    """bugCompat=False stores the OLD MOMENTUM of rejected particles instead of the old
    position (src/HMC.py:176)."""
    g = load_golden("hmc_iso2d_rough")
    po = O.DiagGaussian(g["k"])
    D, P = 2, 64
    q0 = g["z_init"] * float(g["q_std"])
    z, u = g["z"][0], g["u"][0]
    qn, p_fix, acc, _, _ = O.hmc_iter(q0, z, u, g["mass"], float(g["temperature"]), float(g["h"]), int(g["L"]), po,
                                      bug_compat=False)
    assert (~acc).sum() > 0
    ens = E.Ensemble(D, P)
    hmc = E.HMC(ens, float(g["simul_time"]), float(g["h"]), None, potential=E.HarmonicPotential(g["k"]),
                bugCompat=False)
    hmc.integrator.q = q0.copy()
    p_out = np.empty((D, P))
    hmc.step(float(g["temperature"]), p_out=p_out, z=z.copy(), u=u.copy())
    assert rel_err(p_out, p_fix) < 1e-12 and rel_err(hmc.integrator.q, qn) < 1e-12


# ---------------------------------------------------------------------------
# Philox production stream
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_philox_stream_matches_spec(E, dt):
    D, P, seed, it, off = 11, 300, 0x1234567890ABCDEF, 5, 1000
    z = np.empty((D, P), dtype=dt)
    u = np.empty(P, dtype=dt)
    E._lib.philox_fill(E._lib.Context.get(), z, u, seed, it, off)
    zr, ur = O.philox_stream(seed, it, np.arange(P) + off, D, dt)
    assert np.array_equal(u.astype(np.float64), ur)  # integer stream -> exact
    assert np.max(np.abs(z - zr)) < (5e-6 if dt == np.float32 else 1e-13)


@pytest.mark.parametrize("case", ["diag2", "funnel10", "dense20", "dense100"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_production_mode_equals_fed_mode(E, case, dt):
    """z = u = NULL (Philox inside the kernel) == feeding ehmc_philox_fill's output."""
    import torch

    rng = np.random.RandomState(7)
    P = 1000
    if case == "diag2":
        D, pe = 2, E.HarmonicPotential([1.0, 2.0])
    elif case == "funnel10":
        D, pe = 10, E.FunnelPotential(10, 3.0)
    else:
        D = 20 if case == "dense20" else 100
        A = rng.standard_normal((D, D))
        pe = E.GaussianPotential(precision=A @ A.T / D + np.eye(D), mean=rng.standard_normal(D))
    tdt = torch.float32 if dt == np.float32 else torch.float64
    q0 = torch.tensor(rng.standard_normal((D, P)), dtype=tdt, device="cuda")
    mass = torch.tensor(rng.uniform(0.5, 2.0, P), dtype=tdt, device="cuda")
    ctx = E._lib.Context.get()
    h = pe.handle(32 if dt == np.float32 else 64, ctx)
    args = E._lib.make_args(0.1, 0.1**2, 7, KB, 1 / KB, seed=99, iteration=3, particle_offset=12345, flags=1)
    qa, pa = q0.clone(), torch.empty_like(q0)
    acc_a = torch.empty(P, dtype=torch.uint8, device="cuda")
    E._lib.hmc_iter(ctx, h, qa, mass, args, p_out=pa, accept=acc_a)
    z, u = torch.empty_like(q0), torch.empty(P, dtype=tdt, device="cuda")
    E._lib.philox_fill(ctx, z, u, 99, 3, 12345)
    qb, pb = q0.clone(), torch.empty_like(q0)
    acc_b = torch.empty(P, dtype=torch.uint8, device="cuda")
    E._lib.hmc_iter(ctx, h, qb, mass, args, p_out=pb, z=z, u=u, accept=acc_b)
    torch.cuda.synchronize()
    assert torch.equal(acc_a, acc_b)
    assert torch.equal(qa, qb) and torch.equal(pa, pb)
    assert 0 < int(acc_a.sum()) <= P


# ---------------------------------------------------------------------------
# sharding / host path / strides / edge sizes
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["diag2", "dense100"])
def test_shard_invariance_and_column_views(E, case):
    """Particles are independent: running column slices [0,a) and [a,P) with particleOffset
    gives exactly the full run (what the multi-GPU sharding relies on); slices are strided
    views (row stride P), exercising the ld != P path."""
    import torch

    rng = np.random.RandomState(11)
    P, a = 777, 300
    if case == "diag2":
        D, pe = 2, E.HarmonicPotential([1.0, 2.0])
    else:
        D = 100
        A = rng.standard_normal((D, D))
        pe = E.GaussianPotential(precision=A @ A.T / D + np.eye(D))
    q0 = torch.tensor(rng.standard_normal((D, P)), dtype=torch.float32, device="cuda")
    mass = torch.ones(P, dtype=torch.float32, device="cuda")
    ctx = E._lib.Context.get()
    h = pe.handle(32, ctx)
    mk = lambda off: E._lib.make_args(0.05, 0.05**2, 10, KB, 1 / KB, seed=5, iteration=9, particle_offset=off)
    full = q0.clone()
    st_full = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
    E._lib.hmc_iter(ctx, h, full, mass, mk(0), stats=st_full)
    parts = q0.clone()
    st = [torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda") for _ in range(2)]
    E._lib.hmc_iter(ctx, h, parts[:, :a], mass[:a], mk(0), stats=st[0])
    E._lib.hmc_iter(ctx, h, parts[:, a:], mass[a:], mk(a), stats=st[1])
    torch.cuda.synchronize()
    assert torch.equal(full, parts)
    # the statistics are plain sums over particles -> shard sums add up (allreduce semantics).  float32 state: the
    # sums over each warp's 32 particles are float32 (k_small.cuh WarpStats; the dense kernels sum in float64), and a
    # shard boundary that is not a multiple of 32 regroups them
    assert torch.allclose(st_full, st[0] + st[1], rtol=2e-6, atol=1e-4)
    assert st_full[0].item() == st[0][0].item() + st[1][0].item()
    qf = full.double().cpu().numpy()
    assert st_full[0].item() <= P
    np.testing.assert_allclose(st_full[3:3 + D].cpu().numpy(), qf.sum(1), rtol=2e-6, atol=1e-3)
    np.testing.assert_allclose(st_full[3 + D:].cpu().numpy(), (qf * qf).sum(1), rtol=2e-6)


@pytest.mark.parametrize("case", ["diag2", "dense100"])
def test_host_path_equals_device_path(E, case):
    import torch

    rng = np.random.RandomState(13)
    P = 5000
    if case == "diag2":
        D, pe = 2, E.HarmonicPotential([1.0, 2.0])
    else:
        D = 100
        A = rng.standard_normal((D, D))
        pe = E.GaussianPotential(precision=A @ A.T / D + np.eye(D))
    q0 = rng.standard_normal((D, P)).astype(np.float32)
    mass = rng.uniform(0.5, 2, P).astype(np.float32)
    ctx = E._lib.Context.get()
    h = pe.handle(32, ctx)
    args = E._lib.make_args(0.05, 0.05**2, 10, KB, 1 / KB, seed=5, iteration=2)
    qh, ph = q0.copy(), np.empty_like(q0)
    acc_h = np.empty(P, dtype=np.uint8)
    st_h = np.zeros(2 * D + 3)
    E._lib.hmc_iter(ctx, h, qh, mass, args, p_out=ph, accept=acc_h, stats=st_h)
    qd = torch.tensor(q0, device="cuda")
    pd = torch.empty_like(qd)
    acc_d = torch.empty(P, dtype=torch.uint8, device="cuda")
    st_d = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
    E._lib.hmc_iter(ctx, h, qd, torch.tensor(mass, device="cuda"), args, p_out=pd, accept=acc_d, stats=st_d)
    torch.cuda.synchronize()
    assert np.array_equal(qh, qd.cpu().numpy()) and np.array_equal(ph, pd.cpu().numpy())
    assert np.array_equal(acc_h, acc_d.cpu().numpy())
    np.testing.assert_allclose(st_h, st_d.cpu().numpy(), rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("P", [1, 31, 129, 1000])
@pytest.mark.parametrize("case", ["diag3", "dense40"])
def test_ragged_particle_counts_vs_oracle(E, P, case):
    rng = np.random.RandomState(P)
    if case == "diag3":
        D = 3
        k = np.array([2.0, 3.0, 4.0])
        pe, po = E.HarmonicPotential(k), O.DiagGaussian(k)
    else:
        D = 40
        A = rng.standard_normal((D, D))
        prec = A @ A.T / D + np.eye(D)
        mu = rng.standard_normal(D)
        pe, po = E.GaussianPotential(precision=prec, mean=mu), O.DenseGaussian(prec, mu)
    q0 = rng.standard_normal((D, P))
    z = rng.standard_normal((D, P))
    u = rng.uniform(size=P)
    mass = rng.uniform(0.5, 2, P)
    qr, pr, accr, oh, nh = O.hmc_iter(q0, z, u, mass, 1 / KB, 0.2, 9, po)
    ens = E.Ensemble(D, P)
    ens.mass = mass
    hmc = E.HMC(ens, 9 * 0.2 + 1e-9, 0.2, None, potential=pe)
    hmc.integrator.q = q0.copy()
    p_out = np.empty((D, P))
    acc = np.empty(P, dtype=np.uint8)
    hmc.step(1 / KB, p_out=p_out, accept=acc, z=z, u=u)
    assert np.array_equal(acc.astype(bool), accr)
    assert rel_err(hmc.integrator.q, qr) < 1e-12 and rel_err(p_out, pr) < 1e-12


def test_zero_steps_and_empty(E):
    ens = E.Ensemble(2, 8)
    ens.q[:] = 1.5
    ens.p[:] = -0.5
    q, p = E.Leapfrog(ens, 0.1, 0.05, E.HarmonicPotential([1.0, 1.0])).integrate()  # int(0.05/0.1) == 0
    assert np.all(q == 1.5) and np.all(p == -0.5)
    ens0 = E.Ensemble(2, 0)
    q, p = E.Leapfrog(ens0, 0.1, 1.0, E.HarmonicPotential([1.0, 1.0])).integrate()
    assert q.shape == (2, 0)


def test_error_behaviour(E):
    ens = E.Ensemble(2, 4)
    with pytest.raises(ValueError, match="Invalid integration method"):
        E.HMC(ens, 1.0, 0.1, None, potential=E.HarmonicPotential([1.0, 1.0]), method="Euler")
    with pytest.raises(TypeError, match="no CPU fallback"):
        E.HMC(ens, 1.0, 0.1, None, potential=lambda q: 0.5 * np.dot(q, q))
    with pytest.raises(IndexError):
        ens.particle(5)
    with pytest.raises(NotImplementedError):
        E.Integrator(ens, 0.1, 1.0, E.HarmonicPotential([1.0, 1.0])).integrate()
    with pytest.raises(ValueError):
        E.Leapfrog(ens, 0.1, 1.0, E.HarmonicPotential([1.0, 1.0, 1.0]))
    # C-ABI validation: mismatched dtype
    ctx = E._lib.Context.get()
    h = E.HarmonicPotential([1.0, 1.0]).handle(64, ctx)
    with pytest.raises(E._lib.EhmcError, match="float"):
        E._lib.leapfrog(ctx, h, np.zeros((2, 4), np.float32), np.zeros((2, 4), np.float32), np.ones(4, np.float32),
                        0.1, 0.01, 3)
    with pytest.raises(E._lib.EhmcError, match="stride"):
        E._lib.leapfrog(ctx, h, np.zeros((4, 2)).T, np.zeros((2, 4)), np.ones(4), 0.1, 0.01, 3)


# ---------------------------------------------------------------------------
# BASELINE sizes: size-independent properties + spot parity on a particle subset
# ---------------------------------------------------------------------------
def _subset_parity(E, pe, po, D, P, L, h, dt, seed, tol):
    import torch

    tdt = torch.float32 if dt == np.float32 else torch.float64
    gen = torch.Generator(device="cuda").manual_seed(seed)
    q0 = torch.randn((D, P), dtype=tdt, device="cuda", generator=gen)
    z = torch.randn((D, P), dtype=tdt, device="cuda", generator=gen)
    u = torch.rand(P, dtype=tdt, device="cuda", generator=gen)
    mass = torch.ones(P, dtype=tdt, device="cuda")
    ctx = E._lib.Context.get()
    args = E._lib.make_args(h, h**2, L, KB, 1 / KB, flags=1)
    q = q0.clone()
    p = torch.empty_like(q)
    acc = torch.empty(P, dtype=torch.uint8, device="cuda")
    stats = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
    E._lib.hmc_iter(ctx, pe.handle(32 if dt == np.float32 else 64, ctx), q, mass, args, p_out=p, z=z, u=u, accept=acc,
                    stats=stats)
    torch.cuda.synchronize()
    idx = torch.randperm(P, generator=torch.Generator().manual_seed(seed))[:192].sort().values.cuda()
    sub = lambda t: t[..., idx].double().cpu().numpy()
    qr, pr, accr, oh, nh = O.hmc_iter(sub(q0), sub(z), sub(u), np.ones(len(idx)), 1 / KB, h, L, po)
    with np.errstate(over="ignore"):
        clear = np.abs(sub(u) - np.minimum(1, np.exp(oh - nh))) > TIE[dt]
    a = sub(acc).astype(bool)
    assert np.array_equal(a[clear], accr[clear])
    same = a == accr
    assert rel_err(sub(q)[:, same], qr[:, same]) < tol
    assert rel_err(sub(p)[:, same], pr[:, same]) < tol
    n_acc = int(acc.sum().item())
    assert stats[0].item() == n_acc and 0.2 * P < n_acc <= P
    return q0, q, acc


def test_config2_full_size_dense100(E):
    """BASELINE config 2: D = 100, P = 2^20, L = 50 (float32).  Spot parity of 192 random
    particles against the oracle + acceptance statistics."""
    rng = np.random.RandomState(20221018)
    A = rng.standard_normal((100, 100))
    prec = A @ A.T / 100 + np.eye(100)
    pe, po = E.GaussianPotential(precision=prec), O.DenseGaussian(prec)
    _subset_parity(E, pe, po, 100, 1 << 20, 50, 0.05, np.float32, 1, 1e-5)


def test_config5_full_size_funnel(E):
    """BASELINE config 5: Neal's funnel, D = 10, P = 2^22, L = 20 (float32)."""
    pe, po = E.FunnelPotential(10, 3.0), O.Funnel(10, 3.0)
    _subset_parity(E, pe, po, 10, 1 << 22, 20, 0.02, np.float32, 2, 1e-5)


def test_config1_full_run(E):
    """BASELINE config 1 exactly: D = 2, P = 1024, L = 20, 1000 iterations, float64, fed the
    reference's MT19937 stream; compared with the oracle chain (which is bit-equal to the
    reference on the golden prefix)."""
    np.random.seed(20221018)
    ens = E.Ensemble(2, 1024)
    hmc = E.HMC(ens, 1.0, 0.05, None, potential=E.HarmonicPotential([1.0, 1.0]))
    samples, momenta = hmc.getSamples(1000, 1 / KB, 1.0)
    np.random.seed(20221018)
    s_ref, m_ref = O.get_samples(2, 1024, np.ones(1024), O.DiagGaussian([1.0, 1.0]), 1000, 1 / KB, 1.0, 1.0, 0.05)
    assert rel_err(samples, s_ref) < 1e-12 and rel_err(momenta, m_ref) < 1e-12
    # the target is N(0, I): ensemble moments after burn-in
    x = samples[:, :, 200:]
    assert abs(x.mean()) < 0.02 and abs(x.var() - 1.0) < 0.03


@pytest.mark.parametrize("case", ["diag2", "dense100", "funnel10"])
def test_reversibility_property(E, case):
    """Leapfrog is time reversible: integrate, flip p, integrate again -> back at the start.
    A size-independent property checked at 2^18 particles on the device path."""
    import torch

    rng = np.random.RandomState(3)
    P = 1 << 18
    if case == "diag2":
        D, pe = 2, E.HarmonicPotential([1.0, 2.0])
    elif case == "funnel10":
        D, pe = 10, E.FunnelPotential(10, 3.0)
    else:
        D = 100
        A = rng.standard_normal((D, D))
        pe = E.GaussianPotential(precision=A @ A.T / D + np.eye(D))
    ens = E.Ensemble(D, P, dtype=np.float64, device="cuda")
    ens.q.normal_(generator=torch.Generator(device="cuda").manual_seed(1))
    ens.p.normal_(generator=torch.Generator(device="cuda").manual_seed(2))
    q0, p0 = ens.q.clone(), ens.p.clone()
    integ = E.Leapfrog(ens, 0.02, 0.2, pe)
    integ.integrate()
    assert not torch.allclose(ens.q, q0)
    ens.p.neg_()
    integ.integrate()
    ens.p.neg_()
    assert (ens.q - q0).abs().max().item() < 1e-10 * max(1.0, q0.abs().max().item())
    assert (ens.p - p0).abs().max().item() < 1e-10 * max(1.0, p0.abs().max().item())


def test_reversibility_and_energy_tensor_core_path_full_size(E):
    """The float32 tensor-core kernel (3xFP16 split, k_dense_tc3) at config 2's full shape: the
    trajectory is reversible to float32 accuracy and the Hamiltonian error of an L = 50 trajectory is
    the O(h^2) leapfrog error, not operand-rounding drift (acceptance stays high)."""
    import torch

    D, P, L, h = 100, 1 << 20, 50, 0.05
    rng = np.random.RandomState(5)
    A = rng.standard_normal((D, D))
    prec = A @ A.T / D + np.eye(D)
    pe = E.GaussianPotential(precision=prec)
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda")
    ens.q.normal_(generator=torch.Generator(device="cuda").manual_seed(1))
    ens.p.normal_(generator=torch.Generator(device="cuda").manual_seed(2))
    q0, p0 = ens.q.clone(), ens.p.clone()
    lam = torch.tensor(prec, dtype=torch.float64, device="cuda")

    def hamiltonian(q, p, cols):
        qd, pd = q[:, cols].double(), p[:, cols].double()
        return 0.5 * (pd * pd).sum(0) + 0.5 * (qd * (lam @ qd)).sum(0)

    cols = torch.arange(0, P, 257, device="cuda")
    h0 = hamiltonian(ens.q, ens.p, cols)
    integ = E.Leapfrog(ens, h, L * h + 1e-9, pe)
    assert integ.numSteps == L
    integ.integrate()
    h1 = hamiltonian(ens.q, ens.p, cols)
    dh = (h1 - h0).abs()
    assert dh.max().item() < 0.5 and dh.mean().item() < 0.1  # H ~ 100; leapfrog error at h = 0.05
    assert not torch.allclose(ens.q, q0)
    ens.p.neg_()
    integ.integrate()
    ens.p.neg_()
    scale_q, scale_p = q0.abs().max().item(), p0.abs().max().item()
    assert (ens.q - q0).abs().max().item() < 2e-5 * scale_q
    assert (ens.p - p0).abs().max().item() < 2e-5 * scale_p


# ---------------------------------------------------------------------------
# N-body and logistic families
# ---------------------------------------------------------------------------
def test_getAccelNBody_matches_reference(E):
    """src/potential.py:30-53 through the CUDA N-body family, vs the reference's own output."""
    g = load_golden("known_answers")
    q, m = g["nbody_q"], g["nbody_m"]
    for i in range(q.shape[1]):
        a = E.potential.getAccelNBody(q, m, i)
        assert rel_err(a, g["nbody_acc"][:, i]) < 1e-12
    # reference-signed potentials
    assert rel_err(E.potential.nBodyPotential(q, m), g["nbody_pot"]) < 1e-12
    assert rel_err(E.potential.gravitationalPotential(q[:, 0], q[:, 1], m[0], m[1]), g["grav_pot_01"]) < 1e-12


@pytest.mark.parametrize("B,eps", [(5, 0.0), (64, 0.05), (300, 0.0)])
def test_nbody_eval_matches_oracle(E, B, eps):
    rng = np.random.RandomState(B)
    m = rng.uniform(0.5, 1.5, B) / B
    pe, po = E.NBodyPotential(m, G=1.0, eps=eps), O.NBody(m, 1.0, eps)
    q = rng.standard_normal((3 * B, 5))
    for dt in (np.float64, np.float32):
        assert rel_err(pe(q.astype(dt)), po.energy(q)) < 20 * RTOL[dt]
        assert rel_err(pe.gradient(q.astype(dt)), po.grad(q)) < 20 * RTOL[dt]


@pytest.mark.parametrize("N,D", [(40, 6), (1000, 256), (333, 37)])
def test_logistic_eval_matches_oracle(E, N, D):
    rng = np.random.RandomState(N)
    X = rng.standard_normal((N, D)) / np.sqrt(D)
    y = (rng.uniform(size=N) < 0.5).astype(np.float64)
    pe, po = E.LogisticPotential(X, y, 2.0), O.Logistic(X, y, 2.0)
    q = rng.standard_normal((D, 70))
    for dt in (np.float64, np.float32):
        assert rel_err(pe(q.astype(dt)), po.energy(q)) < 20 * RTOL[dt]
        assert rel_err(pe.gradient(q.astype(dt)), po.grad(q)) < 20 * RTOL[dt]


def test_coin_toss_family_KA6_and_hmc(E):
    """The reference's NumPyro sample model (samples/NumpyroExamples/CoinToss): gradient zero at the
    reference biases (KA6), energy/gradient and one HMC iteration against the oracle."""
    import torch

    c1, c2 = [1, 0] * 10, [1] * 15 + [0] * 5
    pe = E.CoinTossPotential.fromObservations(c1, c2)
    po = O.CoinToss([10, 15], [20, 20])
    assert np.array_equal(pe.gradient(np.array([0.5, 0.75])), np.zeros(2))  # KA6, exact in binary
    rng = np.random.RandomState(4)
    P = 500
    q0 = rng.uniform(0.2, 0.8, (2, P))
    for dt in (np.float64, np.float32):
        assert rel_err(pe(q0.astype(dt)), po.energy(q0)) < 20 * RTOL[dt]
        assert rel_err(pe.gradient(q0.astype(dt)), po.grad(q0)) < 20 * RTOL[dt]
    z = rng.standard_normal((2, P)) * 0.3
    u = rng.uniform(size=P)
    mass = rng.uniform(0.5, 2.0, P)
    h, L = 0.01, 12
    qr, pr, accr, oh, nh = O.hmc_iter(q0, z, u, mass, 1 / KB, h, L, po)
    ctx = E._lib.Context.get()
    for dt, bits, tol in ((np.float64, 64, 1e-12), (np.float32, 32, 1e-5)):
        tdt = torch.float64 if bits == 64 else torch.float32
        q = torch.tensor(q0, dtype=tdt, device="cuda")
        p = torch.empty_like(q)
        acc = torch.empty(P, dtype=torch.uint8, device="cuda")
        args = E._lib.make_args(h, h**2, L, KB, 1 / KB, flags=0)
        E._lib.hmc_iter(ctx, pe.handle(bits, ctx), q, torch.tensor(mass, dtype=tdt, device="cuda"), args, p_out=p,
                        z=torch.tensor(z, dtype=tdt, device="cuda"), u=torch.tensor(u, dtype=tdt, device="cuda"),
                        accept=acc)
        torch.cuda.synchronize()
        a = acc.cpu().numpy().astype(bool)
        with np.errstate(over="ignore", invalid="ignore"):
            clear = np.abs(u - np.minimum(1, np.exp(oh - nh))) > TIE[dt]
        clear &= np.isfinite(nh)  # a trajectory that left (0, 1) has H = NaN on both sides
        assert np.array_equal(a[clear], accr[clear])
        same = (a == accr) & np.isfinite(nh)
        assert same.sum() > 0.8 * P
        assert rel_err(q.cpu().numpy()[:, same], qr[:, same]) < tol


@pytest.mark.parametrize("family", ["nbody", "logistic"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_families_production_equals_fed_and_stats(E, family, dt):
    import torch

    rng = np.random.RandomState(17)
    if family == "nbody":
        B, P = 32, 40
        D = 3 * B
        pe = E.NBodyPotential(np.ones(B) / B, G=1.0, eps=0.05)
        h, L = 0.01, 5
    else:
        N, D, P = 200, 24, 300
        X = rng.standard_normal((N, D)) / np.sqrt(D)
        y = (rng.uniform(size=N) < 0.5).astype(np.float64)
        pe = E.LogisticPotential(X, y, 1.0)
        h, L = 0.1, 5
    tdt = torch.float32 if dt == np.float32 else torch.float64
    q0 = torch.tensor(rng.standard_normal((D, P)), dtype=tdt, device="cuda")
    mass = torch.tensor(rng.uniform(0.5, 2.0, P), dtype=tdt, device="cuda")
    ctx = E._lib.Context.get()
    hd = pe.handle(32 if dt == np.float32 else 64, ctx)
    args = E._lib.make_args(h, h**2, L, KB, 1 / KB, seed=3, iteration=11, particle_offset=7, flags=0)
    qa, pa = q0.clone(), torch.empty_like(q0)
    acc_a = torch.empty(P, dtype=torch.uint8, device="cuda")
    st = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
    E._lib.hmc_iter(ctx, hd, qa, mass, args, p_out=pa, accept=acc_a, stats=st)
    z, u = torch.empty_like(q0), torch.empty(P, dtype=tdt, device="cuda")
    E._lib.philox_fill(ctx, z, u, 3, 11, 7)
    qb, pb = q0.clone(), torch.empty_like(q0)
    acc_b = torch.empty(P, dtype=torch.uint8, device="cuda")
    E._lib.hmc_iter(ctx, hd, qb, mass, args, p_out=pb, z=z, u=u, accept=acc_b)
    torch.cuda.synchronize()
    assert torch.equal(acc_a, acc_b) and torch.equal(qa, qb) and torch.equal(pa, pb)
    qf = qa.double().cpu().numpy()
    assert st[0].item() == int(acc_a.sum().item())
    np.testing.assert_allclose(st[3:3 + D].cpu().numpy(), qf.sum(1), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(st[3 + D:].cpu().numpy(), (qf * qf).sum(1), rtol=1e-9)
    # host path gives the same answer
    qh = q0.cpu().numpy().copy()
    sth = np.zeros(2 * D + 3)
    acc_h = np.empty(P, dtype=np.uint8)
    E._lib.hmc_iter(ctx, hd, qh, mass.cpu().numpy(), args, accept=acc_h, stats=sth)
    assert np.array_equal(qh, qa.cpu().numpy()) and np.array_equal(acc_h, acc_a.cpu().numpy())
    np.testing.assert_allclose(sth, st.cpu().numpy(), rtol=1e-12, atol=1e-9)


def test_config4_shape_nbody_subset(E):
    """BASELINE config 4 shape: 4096 bodies x 3-D per particle (D = 12288), Plummer eps = 0.05,
    L = 10, float32; 24 ensemble particles, 2 of them checked against the float64 oracle."""
    import torch

    B, P, L, h = 4096, 24, 10, 0.01
    rng = np.random.RandomState(4)
    m = np.ones(B) / B
    pe, po = E.NBodyPotential(m, G=1.0, eps=0.05), O.NBody(m, 1.0, 0.05)
    q0 = rng.standard_normal((3 * B, P))
    z = rng.standard_normal((3 * B, P))
    u = rng.uniform(size=P)
    ens = E.Ensemble(3 * B, P, dtype=np.float32, device="cuda")
    ens.q.copy_(torch.tensor(q0, dtype=torch.float32))
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pe)
    acc = torch.empty(P, dtype=torch.uint8, device="cuda")
    hmc.step(1 / KB, accept=acc, z=torch.tensor(z, dtype=torch.float32, device="cuda"),
             u=torch.tensor(u, dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    sel = [0, P - 1]
    qr, pr, accr, oh, nh = O.hmc_iter(q0[:, sel], z[:, sel], u[sel], np.ones(2), 1 / KB, h, L, po)
    a = acc.cpu().numpy().astype(bool)[sel]
    clear = np.abs(u[sel] - np.minimum(1, np.exp(oh - nh))) > 2e-3
    assert np.array_equal(a[clear], accr[clear])
    same = a == accr
    assert rel_err(ens.q.cpu().numpy()[:, sel][:, same], qr[:, same]) < 1e-5


def test_config3_shape_logistic_reduced(E):
    """BASELINE config 3 shape reduced in N and P: X 4096 x 256, 512 particles, L = 10, float32."""
    import torch

    N, D, P, L, h = 4096, 256, 512, 10, 0.02
    rng = np.random.RandomState(3)
    X = rng.standard_normal((N, D)) / np.sqrt(D)
    th = rng.standard_normal(D)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-X @ th))).astype(np.float64)
    pe, po = E.LogisticPotential(X, y, 1.0), O.Logistic(X, y, 1.0)
    q0 = rng.standard_normal((D, P))
    z = rng.standard_normal((D, P))
    u = rng.uniform(size=P)
    qr, pr, accr, oh, nh = O.hmc_iter(q0, z, u, np.ones(P), 1 / KB, h, L, po)
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda")
    ens.q.copy_(torch.tensor(q0, dtype=torch.float32))
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pe)
    acc = torch.empty(P, dtype=torch.uint8, device="cuda")
    hmc.step(1 / KB, accept=acc, z=torch.tensor(z, dtype=torch.float32, device="cuda"),
             u=torch.tensor(u, dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    a = acc.cpu().numpy().astype(bool)
    clear = np.abs(u - np.minimum(1, np.exp(oh - nh))) > 2e-3
    assert np.array_equal(a[clear], accr[clear])
    same = a == accr
    assert rel_err(ens.q.cpu().numpy()[:, same], qr[:, same]) < 1e-5


# ---------------------------------------------------------------------------
# dense Gaussian: tensor-core (3xFP16, tcgen05) path vs CUDA-core FP32 path vs oracle
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("D,P,L", [(20, 300, 7), (40, 129, 9), (100, 1000, 50), (104, 257, 12), (17, 64, 0), (56, 200, 3),
                                   (112, 300, 10), (128, 513, 25), (100, 40000, 5)])
def test_dense_tensor_core_path_vs_cuda_core_path_vs_oracle(E, D, P, L):
    import torch

    rng = np.random.RandomState(D)
    A = rng.standard_normal((D, D))
    prec = A @ A.T / D + np.eye(D)
    mu = rng.standard_normal(D)
    pe, po = E.GaussianPotential(precision=prec, mean=mu), O.DenseGaussian(prec, mu)
    q0 = rng.standard_normal((D, P)) + mu[:, None]
    z = rng.standard_normal((D, P))
    u = rng.uniform(size=P)
    mass = rng.uniform(0.5, 2.0, P)
    h = 0.05
    qr, pr, accr, oh, nh = O.hmc_iter(q0, z, u, mass, 1 / KB, h, L, po)
    ctx = E._lib.Context.get()
    hd = pe.handle(32, ctx)
    args = E._lib.make_args(h, h**2, L, KB, 1 / KB, flags=1)
    out = {}
    try:
        # 1 = CUDA cores, 4 = 3xFP16 persistent tensor-core kernel (the default path, D <= 128)
        paths = (1, 4)
        for path in paths:
            ctx.set_option("dense_path", path)
            q = torch.tensor(q0, dtype=torch.float32, device="cuda")
            p = torch.empty_like(q)
            acc = torch.empty(P, dtype=torch.uint8, device="cuda")
            st = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
            E._lib.hmc_iter(ctx, hd, q, torch.tensor(mass, dtype=torch.float32, device="cuda"), args, p_out=p,
                            z=torch.tensor(z, dtype=torch.float32, device="cuda"),
                            u=torch.tensor(u, dtype=torch.float32, device="cuda"), accept=acc, stats=st)
            torch.cuda.synchronize()
            out[path] = (q.cpu().numpy(), p.cpu().numpy(), acc.cpu().numpy().astype(bool), st.cpu().numpy())
    finally:
        ctx.set_option("dense_path", 0)
    with np.errstate(over="ignore"):
        clear = np.abs(u - np.minimum(1, np.exp(oh - nh))) > TIE[np.float32]
    for path in paths:
        q, p, a, st = out[path]
        assert np.array_equal(a[clear], accr[clear]), f"path {path}"
        same = a == accr
        assert rel_err(q[:, same], qr[:, same]) < 1e-5, f"path {path}"
        assert rel_err(p[:, same], pr[:, same]) < 1e-5, f"path {path}"
        assert st[0] == a.sum()
        np.testing.assert_allclose(st[3:3 + D], q.astype(np.float64).sum(1), rtol=1e-9, atol=1e-9)
    print("D", D, {f"path{k}-vs-oracle": float(rel_err(out[k][0], qr)) for k in paths})


@pytest.mark.parametrize("D,P,L", [(100, 700, 20), (40, 129, 0), (128, 300, 7), (24, 1000, 3)])
def test_dense_tensor_core_stormer_verlet(E, D, P, L):
    """method="Stormer-Verlet" (src/integrator.py:142-163: L + 1 position updates, backward-difference
    momentum) on the tensor-core kernel: one HMC iteration and a bare integrate() against the oracle."""
    import torch

    rng = np.random.RandomState(D + L)
    A = rng.standard_normal((D, D))
    prec = A @ A.T / D + np.eye(D)
    mu = rng.standard_normal(D)
    pe, po = E.GaussianPotential(precision=prec, mean=mu), O.DenseGaussian(prec, mu)
    q0 = rng.standard_normal((D, P)) + mu[:, None]
    z = rng.standard_normal((D, P))
    u = rng.uniform(size=P)
    mass = rng.uniform(0.5, 2.0, P)
    h = 0.05
    qr, pr, accr, oh, nh = O.hmc_iter(q0, z, u, mass, 1 / KB, h, L, po, integrator="Stormer-Verlet", bug_compat=False)
    ctx = E._lib.Context.get()
    hd = pe.handle(32, ctx)
    args = E._lib.make_args(h, h**2, L, KB, 1 / KB, integrator=E._lib.STORMER_VERLET, flags=0)
    out = {}
    try:
        for path in (1, 4):  # CUDA cores, tensor cores
            ctx.set_option("dense_path", path)
            q = torch.tensor(q0, dtype=torch.float32, device="cuda")
            p = torch.empty_like(q)
            acc = torch.empty(P, dtype=torch.uint8, device="cuda")
            E._lib.hmc_iter(ctx, hd, q, torch.tensor(mass, dtype=torch.float32, device="cuda"), args, p_out=p,
                            z=torch.tensor(z, dtype=torch.float32, device="cuda"),
                            u=torch.tensor(u, dtype=torch.float32, device="cuda"), accept=acc)
            torch.cuda.synchronize()
            out[path] = (q.cpu().numpy(), p.cpu().numpy(), acc.cpu().numpy().astype(bool))
    finally:
        ctx.set_option("dense_path", 0)
    with np.errstate(over="ignore"):
        clear = np.abs(u - np.minimum(1, np.exp(oh - nh))) > TIE[np.float32]
    for path in (1, 4):
        q, p, a = out[path]
        assert np.array_equal(a[clear], accr[clear]), f"path {path}"
        same = a == accr
        assert rel_err(q[:, same], qr[:, same]) < 1e-5, f"path {path}"
        # the backward-difference momentum (q_{L+1} - q_L) / h loses log10(|q| / (h |v|)) digits in float32
        assert rel_err(p[:, same], pr[:, same]) < 2e-4, f"path {path}"
    # integrate() without Metropolis
    p0 = rng.standard_normal((D, P))
    qi, pi = O.stormer_verlet(q0, p0, mass, h, L, po.grad)
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda")
    ens.q.copy_(torch.tensor(q0, dtype=torch.float32))
    ens.p.copy_(torch.tensor(p0, dtype=torch.float32))
    ens.mass = torch.tensor(mass, dtype=torch.float32, device="cuda")
    qg, pg = E.StormerVerlet(ens, h, L * h + 1e-9, pe).integrate()
    assert rel_err(qg.cpu().numpy(), qi) < 1e-5 and rel_err(pg.cpu().numpy(), pi) < 2e-4


def test_dense_tensor_core_integrate_only(E):
    """Leapfrog.integrate() (no Metropolis) on the tensor-core path."""
    import torch

    D, P, L, h = 100, 500, 20, 0.05
    rng = np.random.RandomState(8)
    A = rng.standard_normal((D, D))
    prec = A @ A.T / D + np.eye(D)
    q0 = rng.standard_normal((D, P))
    p0 = rng.standard_normal((D, P))
    mass = rng.uniform(0.5, 2.0, P)
    qr, pr = O.leapfrog(q0, p0, mass, h, L, O.DenseGaussian(prec).grad)
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda")
    ens.q.copy_(torch.tensor(q0, dtype=torch.float32))
    ens.p.copy_(torch.tensor(p0, dtype=torch.float32))
    ens.mass = torch.tensor(mass, dtype=torch.float32, device="cuda")
    q, p = E.Leapfrog(ens, h, L * h + 1e-9, E.GaussianPotential(precision=prec)).integrate()
    assert rel_err(q.cpu().numpy(), qr) < 1e-5 and rel_err(p.cpu().numpy(), pr) < 1e-5


@pytest.mark.parametrize("scale", [1e-6, 1.0, 3e4])
def test_dense_tensor_core_scale_robustness(E, scale):
    """The fp16 split operands of k_dense_tc3 are rescaled per particle row by a power of two, so a
    trajectory in tiny or huge units (far outside fp16's exponent range) keeps float32 accuracy.
    The dynamics are linear: scaling q and p by s scales the trajectory by s."""
    import torch

    D, P, L, h = 100, 300, 30, 0.05
    rng = np.random.RandomState(11)
    A = rng.standard_normal((D, D))
    prec = A @ A.T / D + np.eye(D)
    mu = rng.standard_normal(D) * scale
    q0 = rng.standard_normal((D, P)) * scale + mu[:, None]
    q0[:, 0] = mu  # a row that starts exactly at the mean (x = 0)
    p0 = rng.standard_normal((D, P)) * scale
    p0[:, 1] = 0.0  # and one at rest
    mass = rng.uniform(0.5, 2.0, P)
    qr, pr = O.leapfrog(q0, p0, mass, h, L, O.DenseGaussian(prec, mu).grad)
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda")
    ens.q.copy_(torch.tensor(q0, dtype=torch.float32))
    ens.p.copy_(torch.tensor(p0, dtype=torch.float32))
    ens.mass = torch.tensor(mass, dtype=torch.float32, device="cuda")
    q, p = E.Leapfrog(ens, h, L * h + 1e-9, E.GaussianPotential(precision=prec, mean=mu)).integrate()
    # positions relative to the spread around the mean (q - mu), not to |mu|
    assert rel_err(q.cpu().numpy() - mu[:, None], qr - mu[:, None]) < 1e-5
    assert rel_err(p.cpu().numpy(), pr) < 1e-5


def test_dense_tensor_core_zero_step(E):
    """h = 0 is a fixed point of the trajectory (q, p unchanged, every proposal accepted)."""
    import torch

    D, P = 64, 200
    rng = np.random.RandomState(12)
    A = rng.standard_normal((D, D))
    prec = A @ A.T / D + np.eye(D)
    q0 = rng.standard_normal((D, P)).astype(np.float32)
    ctx = E._lib.Context.get()
    hd = E.GaussianPotential(precision=prec).handle(32, ctx)
    args = E._lib.make_args(0.0, 0.0, 5, KB, 1 / KB, flags=0)
    q = torch.tensor(q0, device="cuda")
    acc = torch.empty(P, dtype=torch.uint8, device="cuda")
    E._lib.hmc_iter(ctx, hd, q, torch.ones(P, dtype=torch.float32, device="cuda"), args, accept=acc)
    torch.cuda.synchronize()
    assert acc.cpu().numpy().all()
    np.testing.assert_allclose(q.cpu().numpy(), q0, rtol=3e-7, atol=1e-7)


def test_dense_tensor_core_divergent_trajectory_is_rejected(E):
    """A step size far above the stability limit 2 / sqrt(lambda_max) makes |q| grow by orders of magnitude per
    step.  The fp16 split operands of the tensor-core kernel saturate (instead of turning into inf - inf = NaN) and
    the row is rejected outright: q stays finite and unchanged, exactly like the CUDA-core kernel, which carries the
    trajectory to ~1e38 and rejects it with ratio 0.  With the reference's default flags a NaN ratio would have been
    ACCEPTED (src/HMC.py:168-173) and written into q for good."""
    import torch

    D, P, L = 100, 700, 50
    rng = np.random.RandomState(13)
    A = rng.standard_normal((D, D))
    prec = A @ A.T / D + np.eye(D)
    lam_max = np.linalg.eigvalsh(prec).max()
    h = 2.1 / np.sqrt(lam_max)  # 5 % above the stability limit: the stiffest mode grows 1.9x per step, 5e13 in all
    q0 = rng.standard_normal((D, P)).astype(np.float32)
    ctx = E._lib.Context.get()
    hd = E.GaussianPotential(precision=prec).handle(32, ctx)
    args = E._lib.make_args(h, h * h, L, KB, 1 / KB, seed=5, iteration=0, flags=1)  # reference default: NaN accepted
    res = {}
    try:
        for path in (1, 4):
            ctx.set_option("dense_path", path)
            q = torch.tensor(q0, device="cuda")
            acc = torch.empty(P, dtype=torch.uint8, device="cuda")
            st = torch.zeros(2 * D + 3, dtype=torch.float64, device="cuda")
            E._lib.hmc_iter(ctx, hd, q, torch.ones(P, dtype=torch.float32, device="cuda"), args, accept=acc, stats=st)
            torch.cuda.synchronize()
            res[path] = (q.cpu().numpy(), acc.cpu().numpy(), st.cpu().numpy())
    finally:
        ctx.set_option("dense_path", 0)
    for path in (1, 4):
        q, acc, st = res[path]
        assert np.isfinite(q).all(), f"path {path}"
        assert acc.sum() == 0 and np.array_equal(q, q0), f"path {path}"
        assert np.isfinite(st).all() and st[0] == 0 and st[1] == 0, f"path {path}"
    # integrate() has no Metropolis step: the divergence is reported instead (host-backed call raises)
    ens = E.Ensemble(D, P, dtype=np.float32)
    ens.q[:] = q0
    ens.p[:] = rng.standard_normal((D, P))
    with pytest.raises(FloatingPointError):
        E.Leapfrog(ens, h, L * h + 1e-9, E.GaussianPotential(precision=prec)).integrate()
    assert ctx.overflow_count() == 0  # the counter was consumed by the check


@pytest.mark.parametrize("kind", ["kappa1e6", "span2^30", "moderate"])
def test_dense_ill_conditioned_precision_guard(E, kind):
    """dense_path 0 (auto) on precision matrices that are hard for the fp16 split: entries far below max |Lambda|
    flush in fp16, and coordinate scales spread past the 2^7 headroom of the per-row scale.  ehmc_potential_create
    detects both and routes such potentials to the exact CUDA-core kernel.  float32 state itself limits what any
    kernel can reach here (Lambda dx with |dx| = 6e-8 |x| is kappa^(1/2) times the gradient's own rounding), so
    the bar is: the default path is within 1e-5, or within 4x of the exact float32 FMA kernel's own error."""
    import torch

    D, P, L = 64, 400, 20
    rng = np.random.RandomState(21)
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    if kind == "kappa1e6":
        prec = (Q * np.logspace(0, 6, D)) @ Q.T
    elif kind == "span2^30":
        s = np.logspace(0, 4.5, D)  # coordinate scales over 2^15: diag(Lambda) spans 2^30
        A = rng.standard_normal((D, D))
        prec = (A @ A.T / D + np.eye(D)) * np.outer(s, s)
    else:
        prec = (Q * np.logspace(0, 2, D)) @ Q.T  # kappa = 100
    prec = 0.5 * (prec + prec.T)
    sd = 1.0 / np.sqrt(np.diag(prec))
    q0 = rng.standard_normal((D, P)) * sd[:, None]
    z = rng.standard_normal((D, P))
    u = rng.uniform(size=P)
    h = 0.3 / np.sqrt(np.linalg.eigvalsh(prec).max())
    qr, pr, accr, oh, nh = O.hmc_iter(q0, z, u, np.ones(P), 1 / KB, h, L, O.DenseGaussian(prec, np.zeros(D)))
    ctx = E._lib.Context.get()
    hd = E.GaussianPotential(precision=prec).handle(32, ctx)
    args = E._lib.make_args(h, h * h, L, KB, 1 / KB, flags=0)
    with np.errstate(over="ignore"):
        clear = np.abs(u - np.minimum(1, np.exp(oh - nh))) > TIE[np.float32]
    res = {}
    try:
        for path in (0, 1):
            ctx.set_option("dense_path", path)
            q = torch.tensor(q0, dtype=torch.float32, device="cuda")
            p = torch.empty_like(q)
            acc = torch.empty(P, dtype=torch.uint8, device="cuda")
            E._lib.hmc_iter(ctx, hd, q, torch.ones(P, dtype=torch.float32, device="cuda"), args, p_out=p,
                            z=torch.tensor(z, dtype=torch.float32, device="cuda"),
                            u=torch.tensor(u, dtype=torch.float32, device="cuda"), accept=acc)
            torch.cuda.synchronize()
            a = acc.cpu().numpy().astype(bool)
            same = a == accr
            # every coordinate in units of its own scale
            eq = np.abs(q.cpu().numpy()[:, same] - qr[:, same]) / sd[:, None]
            res[path] = (a, float(eq.max() / np.abs(qr[:, same] / sd[:, None]).max()),
                         rel_err(p.cpu().numpy()[:, same], pr[:, same]), q.cpu().numpy())
    finally:
        ctx.set_option("dense_path", 0)
    print(kind, {k: v[1:3] for k, v in res.items()})
    assert np.array_equal(res[0][0][clear], accr[clear])
    assert res[0][1] < max(1e-5, 4 * res[1][1]) and res[0][2] < max(1e-5, 4 * res[1][2])
    if kind == "span2^30":
        assert np.array_equal(res[0][3], res[1][3])  # the guard chose the CUDA-core kernel: identical bits
    if kind == "moderate":
        assert res[0][1] < 1e-5 and res[0][2] < 1e-5


@pytest.mark.parametrize("case", ["diag2", "funnel10", "dense100", "dense40_cuda_cores", "nbody", "logistic", "logistic_tc",
                                  "logistic_tcs"])
def test_no_out_of_bounds_writes_canary(E, case):
    """Every in/out tensor of ehmc_hmc_iter is a window into a larger canary-filled allocation (one guard row
    above and below, 37 guard columns left and right, ragged P): after the call every guard element still
    holds the canary.  (compute-sanitizer is closed on this pool; this is the bounds check of our own.)"""
    import torch

    rng = np.random.RandomState(51)
    P, L, h = 333, 3, 0.05
    path = 0
    if case == "diag2":
        D, pot = 2, E.HarmonicPotential([1.0, 2.0])
    elif case == "funnel10":
        D, pot = 10, E.FunnelPotential(10, 3.0)
    elif case.startswith("dense"):
        D = 100 if case == "dense100" else 40
        A = rng.standard_normal((D, D))
        pot = E.GaussianPotential(precision=A @ A.T / D + np.eye(D), mean=rng.standard_normal(D))
        path = 1 if case.endswith("cuda_cores") else 0
    elif case == "nbody":
        B = 21
        D, pot, h = 3 * B, E.NBodyPotential(np.ones(B) / B, G=1.0, eps=0.1), 0.01
    else:
        N, D = 300, 24
        X = rng.standard_normal((N, D)) / np.sqrt(D)
        y = (rng.uniform(size=N) < 0.5).astype(np.float64)
        pot = E.LogisticPotential(X, y, 1.0, precision={"logistic": "fp32", "logistic_tc": "bf16",
                                                         "logistic_tcs": "fp16x3"}[case])
    CAN, G = -777.25, 37

    def window(rows, dtype=torch.float32):
        big = torch.full((rows + 2, P + 2 * G), CAN, dtype=dtype, device="cuda")
        return big, big[1:rows + 1, G:G + P]

    qb, q = window(D)
    pb, p = window(D)
    q.copy_(torch.tensor(rng.standard_normal((D, P)), dtype=torch.float32))
    mb = torch.full((P + 2 * G,), CAN, dtype=torch.float32, device="cuda")
    mass = mb[G:G + P]
    mass.fill_(1.3)
    ab = torch.full((P + 2 * G,), 77, dtype=torch.uint8, device="cuda")
    acc = ab[G:G + P]
    sb = torch.full((2 * D + 3 + 2 * G,), CAN, dtype=torch.float64, device="cuda")
    st = sb[G:G + 2 * D + 3]
    ctx = E._lib.Context.get()
    ctx.set_option("dense_path", path)
    try:
        args = E._lib.make_args(h, h * h, L, KB, 1 / KB, seed=2, iteration=1)
        E._lib.hmc_iter(ctx, pot.handle(32, ctx), q, mass, args, p_out=p, accept=acc, stats=st)
        torch.cuda.synchronize()
    finally:
        ctx.set_option("dense_path", 0)

    def guards_intact(big, inner_rows, value):
        g = big.clone()
        if big.dim() == 2:
            g[1:inner_rows + 1, G:G + P] = value
        else:
            g[G:G + inner_rows] = value
        return bool((g == value).all().item())

    assert guards_intact(qb, D, CAN) and guards_intact(pb, D, CAN)
    assert guards_intact(mb, P, CAN) and guards_intact(ab, P, 77) and guards_intact(sb, 2 * D + 3, CAN)
    assert torch.isfinite(q).all() and torch.isfinite(p).all() and (acc <= 1).all()
    assert st[0].item() == acc.sum().item()


# ---------------------------------------------------------------------------
# production loop: statistics, adaptation, ESS
# ---------------------------------------------------------------------------
def test_run_adaptation_and_moments(E):
    """HMC.run on a 20-D correlated Gaussian: the adapted step size brings the ensemble
    acceptance to the target, the streaming moments match the target covariance, ESS > 0."""
    import torch

    D, P = 20, 1 << 15
    rng = np.random.RandomState(2)
    A = rng.standard_normal((D, D))
    prec = A @ A.T / D + np.eye(D)
    cov = np.linalg.inv(prec)
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=7)
    ens.setPosition(1.0)
    hmc = E.HMC(ens, 1.5, 0.02, None, potential=E.GaussianPotential(precision=prec), seed=7)
    r = hmc.run(200, 1 / KB, adapt=True, targetAccept=0.8, adaptIterations=160)
    assert abs(np.mean(r["meanAcceptProb"][170:]) - 0.8) < 0.05
    assert hmc.stepSize > 0.05  # grew from the tiny initial value
    r2 = hmc.run(120, 1 / KB, traceParticles=128)
    np.testing.assert_allclose(r2["var"].numpy(), np.diag(cov), rtol=0.06)
    assert np.all(np.abs(r2["mean"].numpy()) < 0.05)
    m, scaled = E.diagnostics.ess_min_over_dims(r2["trace"], numParticlesTotal=P)
    assert 0 < m <= 128 * 120 * 3 and scaled == pytest.approx(m * P / 128)


@pytest.mark.parametrize("case", ["funnel10", "dense40"])
def test_run_device_adaptation_matches_host_adaptation(E, case):
    """deviceAdapt=True (step size + iteration counter in device-resident ehmc_dynamic blocks, update by
    ehmc_adapt_step on a side stream, optionally replayed from a CUDA graph) walks the same chain as the
    host-side loop of HMC.run: same Philox stream, same Robbins-Monro rule, same one-iteration-stale
    consumption of the statistics."""
    import torch

    P, S, L = 4096, 27, 6
    rng = np.random.RandomState(21)
    if case == "funnel10":
        D, pot = 10, E.FunnelPotential(10, 3.0)
    else:
        D = 40
        A = rng.standard_normal((D, D))
        pot = E.GaussianPotential(precision=A @ A.T / D + np.eye(D))
    q0 = torch.tensor(rng.standard_normal((D, P)), dtype=torch.float32, device="cuda")

    def fresh():
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=5)
        ens.q.copy_(q0)
        return ens, E.HMC(ens, L * 0.02 + 1e-9, 0.02, None, potential=pot, seed=5, bugCompat=False)

    ens_h, hmc_h = fresh()
    rh = hmc_h.run(S, 1 / KB, adapt=True, targetAccept=0.8, adaptIterations=20, keepNumSteps=True, fused=False)
    outs = {}
    for graph in (False, True):
        ens_d, hmc_d = fresh()
        r = hmc_d.run(S, 1 / KB, adapt=True, targetAccept=0.8, adaptIterations=20, deviceAdapt=True, graph=graph)
        assert hmc_d.iteration == S and hmc_d.integrator.numSteps == L
        np.testing.assert_allclose(r["stepSize"], rh["stepSize"], rtol=1e-5)
        np.testing.assert_allclose(r["meanAcceptProb"], rh["meanAcceptProb"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(r["meanH"], rh["meanH"], rtol=1e-4, atol=1e-4)
        assert abs(hmc_d.stepSize - hmc_h.stepSize) < 1e-5 * hmc_h.stepSize
        assert r["stepSize"][0] == 0.02 and r["stepSize"][1] == 0.02 and r["stepSize"][2] != 0.02  # one iteration stale
        assert r["stepSize"][22] == r["stepSize"][26]  # adaptation stops after adaptIterations
        np.testing.assert_allclose(r["var"].numpy(), rh["var"].numpy(), rtol=1e-3)
        outs[graph] = ens_d.q.clone()
    # eager and graph-replayed runs are the same sequence of launches: bit-identical ensembles
    assert torch.equal(outs[False], outs[True])
    # and the chain itself follows the host-adapted one (step sizes agree to ~1e-7 relative)
    assert rel_err(outs[True].cpu().numpy(), ens_h.q.cpu().numpy()) < 1e-3


@pytest.mark.parametrize("case", ["funnel10", "diag3", "dense8", "coin2"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("lag", [1, 2])
def test_run_fused_matches_host_loop(E, case, dt, lag):
    """HMC.run(fused=True): trajectory kernel, statistics, step-size update in ONE persistent cooperative launch
    (ehmc_hmc_run_ensemble) against the iteration-by-iteration host loop: same Philox stream, same Robbins-Monro
    schedule (the statistics of iteration k set the step size of iteration k + 1 + adaptLag).  The statistics are
    float64 sums in a different (fixed) order, so the step sizes agree to 1e-12 and the chains to the float rounding
    of h."""
    import torch

    P, S, L = 5000, 31, 5
    rng = np.random.RandomState(22)
    if case == "funnel10":
        D, pot, h0 = 10, E.FunnelPotential(10, 3.0), 0.02
    elif case == "diag3":
        D, pot, h0 = 3, E.HarmonicPotential([1.0, 4.0, 0.25]), 0.1
    elif case == "coin2":
        D, pot, h0 = 2, E.CoinTossPotential([10, 15], [20, 20]), 0.01
    else:
        D = 8
        A = rng.standard_normal((D, D))
        pot, h0 = E.GaussianPotential(precision=A @ A.T / D + np.eye(D), mean=rng.standard_normal(D)), 0.05
    tdt = torch.float32 if dt == np.float32 else torch.float64
    q0 = rng.uniform(0.3, 0.7, (D, P)) if case == "coin2" else rng.standard_normal((D, P))
    q0 = torch.tensor(q0, dtype=tdt, device="cuda")
    ctx = E._lib.Context.get()

    def fresh():
        ens = E.Ensemble(D, P, dtype=dt, device="cuda", seed=5)
        ens.q.copy_(q0)
        return ens, E.HMC(ens, L * h0 + 1e-9, h0, None, potential=pot, seed=5, bugCompat=False)

    ens_h, hmc_h = fresh()
    rh = hmc_h.run(S, 1 / KB, adapt=True, targetAccept=0.8, adaptIterations=24, keepNumSteps=True, fused=False,
                   traceParticles=7, adaptLag=lag)
    ens_f, hmc_f = fresh()
    n0 = ctx.launch_count()
    rf = hmc_f.run(S, 1 / KB, adapt=True, targetAccept=0.8, adaptIterations=24, keepNumSteps=True, traceParticles=7,
                   adaptLag=lag)
    assert rf.get("fused") and ctx.launch_count() - n0 == 1  # the default picks the fused path: ONE launch
    assert hmc_f.iteration == S and hmc_f.integrator.numSteps == L
    np.testing.assert_allclose(rf["stepSize"], rh["stepSize"], rtol=1e-12)
    assert all(rf["stepSize"][k] == h0 for k in range(lag + 1)) and rf["stepSize"][lag + 1] != h0  # `lag` iterations stale
    assert rf["stepSize"][25 + lag] == rf["stepSize"][30]  # adaptation stops after adaptIterations
    assert abs(hmc_f.stepSize - hmc_h.stepSize) <= 1e-12 * hmc_h.stepSize
    np.testing.assert_array_equal(rf["acceptRate"], rh["acceptRate"])
    np.testing.assert_allclose(rf["meanAcceptProb"], rh["meanAcceptProb"], rtol=1e-12)
    np.testing.assert_allclose(rf["meanH"], rh["meanH"], rtol=1e-12, atol=1e-12)
    # (the float64 coin-toss chain is chaotic near the 1 / q poles: the last-bit difference between the device's and
    # libm's exp(log h) reaches the moments of later iterations)
    loose = case == "coin2" and dt == np.float64
    np.testing.assert_allclose(rf["mean"].numpy(), rh["mean"].numpy(), rtol=1e-6 if loose else 1e-11, atol=1e-13)
    np.testing.assert_allclose(rf["var"].numpy(), rh["var"].numpy(), rtol=1e-6 if loose else 1e-10)
    def same_chain(a, b):
        if dt == np.float32:
            return torch.equal(a, b)  # (float) h is the same number: identical bits
        # float64: exp(log h) on the device and in libm differ in the last bit, and the chain follows h exactly
        if case == "coin2":
            return True  # trajectories that leave (0, 1) are chaotic (1 / q poles): only the statistics above are compared
        return bool(np.allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-9, atol=1e-11, equal_nan=True))

    assert same_chain(ens_f.q, ens_h.q)
    assert same_chain(rf["trace"], rh["trace"])
    # a second call continues the chain (Philox iteration counter, adapted step size) without adaptation
    r2h = hmc_h.run(5, 1 / KB, fused=False)
    r2f = hmc_f.run(5, 1 / KB)
    assert same_chain(ens_f.q, ens_h.q)
    np.testing.assert_allclose(r2f["stepSize"], r2h["stepSize"], rtol=1e-12)  # device exp() vs libm: 1 ulp


@pytest.mark.parametrize("lag", [1, 2, 3])
def test_run_fused_two_ranks(E, tmp_path, lag):
    """The in-kernel all-reduce of the fused ensemble run: two processes (sharing this GPU: CUDA IPC mailboxes work
    within one device and the time-sliced persistent kernels still make progress), each with half of the ensemble,
    against one process with all of it.  Philox ids are global, so the ensemble statistics -- and with them the
    adapted step sizes -- agree to the rounding of the float64 sums, and every shard walks the single-process
    chain."""
    import socket
    import subprocess
    import sys

    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fused_rank_worker.py")
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()

    def launch(rank, world, out):
        return subprocess.Popen([sys.executable, worker, "--rank", str(rank), "--world", str(world), "--port", str(port),
                                 "--out", out, "--lag", str(lag)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                text=True)

    single = str(tmp_path / "single.npz")
    p = launch(0, 1, single)
    out, _ = p.communicate(timeout=300)
    assert p.returncode == 0, out
    procs = [launch(r, 2, str(tmp_path / f"rank{r}.npz")) for r in range(2)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for k in procs:
                k.kill()
            pytest.fail("the two-rank fused run did not finish (mailbox all-reduce stuck)")
        outs.append(o)
        assert p.returncode == 0, o
    ref = np.load(single)
    shards = [np.load(str(tmp_path / f"rank{r}.npz")) for r in range(2)]
    for sh in shards:
        np.testing.assert_allclose(sh["stepSize"], ref["stepSize"], rtol=1e-12)
        np.testing.assert_array_equal(sh["acceptRate"], ref["acceptRate"])
        np.testing.assert_allclose(sh["meanH"], ref["meanH"], rtol=1e-12)
        np.testing.assert_allclose(sh["var"], ref["var"], rtol=1e-10)
        assert abs(sh["final"] - ref["final"]) <= 1e-12 * ref["final"]
        np.testing.assert_array_equal(sh["q"], ref["q"][:, int(sh["lo"]):int(sh["hi"])])
    # both ranks hold the SAME reduced numbers bit for bit (rank-ordered sum)
    np.testing.assert_array_equal(shards[0]["stepSize"], shards[1]["stepSize"])
    np.testing.assert_array_equal(shards[0]["meanH"], shards[1]["meanH"])
    np.testing.assert_array_equal(shards[0]["trace"], ref["trace"])


# ---------------------------------------------------------------------------
# mass adaptation (HMC.run(adaptMass=True)): rescaled coordinates == diagonal mass matrix
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["diag3", "dense8", "funnel5"])
def test_rescaled_potential_is_hmc_with_diagonal_mass(E, case):
    """One HMC iteration on Potential.rescaled(s) in the coordinates q / s (what run(adaptMass=True) launches) against
    the oracle's leapfrog with the diagonal mass matrix M_d = mass / s_d^2 written out in the ORIGINAL coordinates:
    same fed z and u, float64, 1e-11."""
    rng = np.random.RandomState(31)
    P, L, h = 257, 7, 0.11
    if case == "diag3":
        D, pe, po = 3, E.HarmonicPotential([4.0, 0.25, 1.0]), O.DiagGaussian(np.array([4.0, 0.25, 1.0]))
        s = np.array([0.5, 2.0, 3.0])
    elif case == "dense8":
        D = 8
        A = rng.standard_normal((D, D))
        prec, mu = A @ A.T / D + np.eye(D), rng.standard_normal(D)
        pe, po = E.GaussianPotential(precision=prec, mean=mu), O.DenseGaussian(prec, mu)
        s = rng.uniform(0.3, 3.0, D)
    else:
        D, pe, po = 5, E.FunnelPotential(5, 3.0), O.Funnel(5, 3.0)
        s = np.array([2.5, 0.7, 0.7, 0.7, 0.7])
        assert np.array_equal(pe.projectScales([2.5, 0.5, 0.7, 0.9, 0.7])[1:], np.full(4, np.sqrt(np.mean(np.array([0.5, 0.7, 0.9, 0.7]) ** 2))))
    q0 = rng.standard_normal((D, P))
    z = rng.standard_normal((D, P))
    u = rng.uniform(size=P)
    mass = rng.uniform(0.5, 2.0, P)
    qr, accr, oh, nh = O.hmc_iter_diag_mass(q0, z, u, mass, 1.0 / s**2, 1 / KB, h, L, po)
    ps = pe.rescaled(s)
    assert type(ps) is type(pe) and ps.family == pe.family
    # the rescaled descriptor evaluates U(s * q') on the GPU
    np.testing.assert_allclose(ps(q0 / s[:, None]), po.energy(q0), rtol=1e-12, atol=1e-12)
    ens = E.Ensemble(D, P)
    ens.q[:] = q0 / s[:, None]
    ens.mass[:] = mass
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=ps, bugCompat=False)
    acc = np.empty(P, dtype=np.uint8)
    hmc.step(1 / KB, accept=acc, z=z, u=u)
    clear = np.abs(u - np.minimum(1, np.exp(oh - nh))) > 1e-9
    assert np.array_equal(acc.astype(bool)[clear], accr[clear])
    same = acc.astype(bool) == accr
    assert rel_err((ens.q * s[:, None])[:, same], qr[:, same]) < 1e-11


@pytest.mark.parametrize("case", ["diag10", "dense20", "funnel10"])
def test_run_adapt_mass(E, case):
    """HMC.run(adaptMass=True): the all-reduced moments become per-dimension masses (coordinate scales) during the
    warm-up; ensemble.q, mean, var and trace come back in the original coordinates; ESS per iteration of the worst
    dimension rises against step-size adaptation alone (same seeds, same number of iterations)."""
    import torch

    from physicsbasedbayesianinference_b200 import diagnostics

    rng = np.random.RandomState(41)
    P, L, warm, S = 8192, 16, 200, 300
    if case == "diag10":
        D = 10
        sd = np.logspace(-1.5, 1.0, D)
        pot, h0, q_sd = E.HarmonicPotential(1.0 / sd**2), 0.01, sd
    elif case == "dense20":
        D = 20
        A = rng.standard_normal((D, D))
        corr = A @ A.T / D + 2.0 * np.eye(D)
        dd = np.sqrt(np.diag(corr))
        corr = corr / dd[:, None] / dd[None, :]
        sd = np.logspace(-1.0, 1.0, D)
        cov = corr * sd[:, None] * sd[None, :]
        pot, h0, q_sd = E.GaussianPotential(cov=cov), 0.01, sd
    else:
        D = 10
        pot, h0, q_sd = E.FunnelPotential(D, 3.0), 0.05, np.ones(D)
    res = {}
    for am in (False, True):
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=3)
        ens.q.copy_(torch.tensor(rng.standard_normal((D, P)) * q_sd[:, None] * 0.5, dtype=torch.float32))
        hmc = E.HMC(ens, L * h0 + 1e-9, h0, None, potential=pot, seed=3, bugCompat=False)
        rw = hmc.run(warm, 1 / KB, adapt=True, adaptMass=am, keepNumSteps=True)
        r = hmc.run(S, 1 / KB, traceParticles=128)
        torch.cuda.synchronize()
        assert torch.isfinite(ens.q).all() and hmc.potential is pot
        ess, _ = diagnostics.ess_min_over_dims(r["trace"].double().cpu())
        res[am] = dict(ess=ess / S, scale=hmc.massScale, var=r["var"].numpy(), q=ens.q, acc=float(np.mean(r["acceptRate"])),
                       h=hmc.stepSize, warm=rw)
    assert res[False]["scale"] is None and res[True]["scale"] is not None
    assert len(res[True]["warm"]["massScale"]) == warm and len(res[True]["warm"]["stepSize"]) == warm
    if case != "funnel10":
        # the adapted scales are the target's marginal standard deviations, and the moments come back in the
        # original coordinates
        np.testing.assert_allclose(res[True]["scale"], sd, rtol=0.15)
        np.testing.assert_allclose(res[True]["var"], sd**2, rtol=0.15)
        np.testing.assert_allclose(res[True]["q"].double().std(dim=1).cpu().numpy(), sd, rtol=0.1)
        assert res[True]["ess"] > 3.0 * res[False]["ess"], (res[True]["ess"], res[False]["ess"])
    else:
        sc = res[True]["scale"]
        assert np.all(sc[1:] == sc[1]) and sc[0] > 1.5  # x dimensions pooled; v has sd 3
        assert res[True]["ess"] > 1.0 * res[False]["ess"], (res[True]["ess"], res[False]["ess"])
    print(f"adaptMass {case}: ESS/iteration (worst dimension) {res[False]['ess']:.4f} -> {res[True]['ess']:.4f}; "
          f"step {res[False]['h']:.4g} -> {res[True]['h']:.4g}; acceptance {res[False]['acc']:.3f} -> {res[True]['acc']:.3f}")


# ---------------------------------------------------------------------------
# logistic regression on the tensor cores (bf16 GEMM chain, tcgen05)
# ---------------------------------------------------------------------------
def _bf16_round(a):
    b = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32).astype(np.float64)


@pytest.mark.parametrize("N,D,P", [(1000, 256, 300), (128, 16, 128), (700, 40, 77), (5000, 256, 1024)])
def test_logistic_tensor_core_gradient(E, N, D, P):
    """k_logistic_tc vs the float64 oracle evaluated on the SAME bf16-rounded X and theta (tight:
    only fp32 accumulation, tanh.approx and the bf16 rounding of the residual differ) and vs the
    exact oracle (loose: the bf16 input rounding itself)."""
    rng = np.random.RandomState(N + D)
    X = rng.standard_normal((N, D)) / np.sqrt(D)
    y = (rng.uniform(size=N) < 0.5).astype(np.float64)
    th = rng.standard_normal((D, P))
    pe = E.LogisticPotential(X, y, 2.0, precision="bf16")
    g = pe.gradient(th.astype(np.float32))
    u = pe(th.astype(np.float32))
    po_b = O.Logistic(_bf16_round(X), y, 2.0)
    thb = _bf16_round(th)
    # prior term uses the unrounded theta in the kernel
    g_ref = po_b.grad(thb) - thb / 4.0 + th / 4.0
    u_ref = po_b.energy(thb) - 0.5 * np.sum(thb * thb, 0) / 4.0 + 0.5 * np.sum(th * th, 0) / 4.0
    assert rel_err(g, g_ref) < 5e-3  # residual rounded to bf16 before the second GEMM
    assert rel_err(u, u_ref) < 1e-4
    po = O.Logistic(X, y, 2.0)
    assert rel_err(g, po.grad(th)) < 2e-2
    assert rel_err(u, po.energy(th)) < 2e-3


@pytest.mark.parametrize("N,D,P", [(1000, 256, 300), (128, 16, 128), (700, 40, 77), (5000, 256, 1024), (64, 3, 5),
                                   (4096, 200, 513)])
def test_logistic_split_tensor_core_gradient(E, N, D, P):
    """k_logistic_tcs (tcgen05 GEMM chain, 3-pass fp16 split of X, theta and the residual) against the float64
    oracle on the UNROUNDED inputs: float32-level accuracy, three orders of magnitude tighter than the bf16 chain."""
    rng = np.random.RandomState(N + D)
    X = rng.standard_normal((N, D)) / np.sqrt(D)
    X[:, 0] *= 30.0  # a column on another scale (the guard of "auto" would still accept it)
    y = (rng.uniform(size=N) < 0.5).astype(np.float64)
    th = rng.standard_normal((D, P))
    th[:, 0] *= 1e-3  # rows on very different scales: the per-row power-of-two scale
    th[:, 1] *= 1e3
    if P > 2:
        th[:, 2] = 0.0
    pe, po = E.LogisticPotential(X, y, 2.0, precision="fp16x3"), O.Logistic(X, y, 2.0)
    g = pe.gradient(th.astype(np.float32))
    u = pe(th.astype(np.float32))
    th32 = th.astype(np.float32).astype(np.float64)
    g_ref, u_ref = po.grad(th32), po.energy(th32)
    # per particle: gradient error relative to that particle's largest gradient component
    eg = np.max(np.abs(g - g_ref), axis=0) / np.max(np.abs(g_ref), axis=0)
    eu = np.abs(u - u_ref) / np.abs(u_ref)
    print("split gradient", N, D, P, "max rel grad err", eg.max(), "energy", eu.max())
    # (the particle scaled by 1e3 has logits ~ 1e3: their float32 rounding alone moves sigmoid by 1e-5 near s = 0)
    assert np.median(eg) < 1.5e-6 and eg.max() < 1e-5
    assert eu.max() < 1e-6  # float32 output rounding (6e-8) plus the MUFU lg2 / ex2 of softplus


def test_logistic_precision_auto_guard(E):
    """precision="auto": the tensor-core split when X is representable, the exact CUDA-core kernel when a column sits
    2^20 below the largest entry (its lo parts would be fp16 subnormals)."""
    rng = np.random.RandomState(77)
    N, D, P = 500, 32, 200
    X = rng.standard_normal((N, D)) / np.sqrt(D)
    y = (rng.uniform(size=N) < 0.5).astype(np.float64)
    th = rng.standard_normal((D, P)).astype(np.float32)
    ctx = E._lib.Context.get()
    for bad in (False, True):
        Xc = X.copy()
        if bad:
            Xc[:, 5] *= 2.0 ** -20
            th[5] *= 2.0 ** 20
        auto, exact = E.LogisticPotential(Xc, y, 1.0), E.LogisticPotential(Xc, y, 1.0, precision="fp32")
        ga, ge = auto.gradient(th), exact.gradient(th)
        ref = O.Logistic(Xc, y, 1.0).grad(th.astype(np.float64))
        if bad:
            assert np.array_equal(ga, ge)  # fell back: same kernel, same bits
        else:
            assert not np.array_equal(ga, ge)  # tensor cores: different rounding
        scale = np.max(np.abs(ref), axis=1, keepdims=True)
        assert np.max(np.abs(ga - ref) / scale) < 1e-5


def test_config4_full_size(E):
    """BASELINE config 4 at FULL size (4096 bodies x 3-D, ensemble of 1024, L = 10): one HMC iteration of the whole
    ensemble with fed momenta and uniforms; 8 of the 1024 particles against the float64 oracle, whose results are the
    committed fixture tests/golden/c4_full_8.npz (tests/golden/make_golden_c4.py: the oracle needs 13 s per particle)."""
    import torch

    from tests.golden.make_golden_c4 import B, EPS, H, L, P, config4_inputs

    g = np.load(os.path.join(GOLDEN, "c4_full_8.npz"))
    sel = g["sel"]
    q0, z, u = config4_inputs()
    m = np.ones(B) / B
    ens = E.Ensemble(3 * B, P, dtype=np.float32, device="cuda")
    ens.q.copy_(torch.tensor(q0, dtype=torch.float32))
    hmc = E.HMC(ens, L * H + 1e-9, H, None, potential=E.NBodyPotential(m, G=1.0, eps=EPS))
    assert hmc.integrator.numSteps == L
    acc = torch.empty(P, dtype=torch.uint8, device="cuda")
    hmc.step(1 / KB, accept=acc, z=torch.tensor(z, dtype=torch.float32, device="cuda"),
             u=torch.tensor(u, dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    a = acc.cpu().numpy().astype(bool)[sel]
    with np.errstate(over="ignore"):
        clear = np.abs(u[sel] - np.minimum(1, np.exp(g["oldH"] - g["newH"]))) > TIE[np.float32]
    assert np.array_equal(a[clear], g["accept"][clear])
    same = a == g["accept"]
    assert same.sum() >= 6
    qg = ens.q.cpu().numpy()[:, sel]
    err = float(np.max(np.abs(qg[:, same] - g["q1"][:, same])) / float(g["q1_absmax"]))
    print("config 4 full size: rel err q", err, "accepted", int(a.sum()), "of", len(sel))
    assert err < 1e-5


def test_config3_full_size(E):
    """BASELINE config 3 at FULL size on the tensor-core path: X 100 000 x 256, 65 536 particles, L = 10, float32
    state, fed momenta and uniforms; 128 random particles against the float64 oracle at the north_star's 1e-5.
    Positions start around the generating parameter (the posterior's neighbourhood: sd ~ 0.1)."""
    import torch

    N, D, P, L, h = 100_000, 256, 65_536, 10, 0.1  # h omega_max = 1.04: 107 of the 128 oracle proposals are accepted
    rng = np.random.RandomState(33)
    X = rng.standard_normal((N, D)) / np.sqrt(D)
    th_true = rng.standard_normal(D)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-X @ th_true))).astype(np.float64)
    pe = E.LogisticPotential(X, y, 1.0, precision="fp16x3")
    q0 = (th_true[:, None] + 0.1 * rng.standard_normal((D, P))).astype(np.float32)
    z = rng.standard_normal((D, P)).astype(np.float32)
    u = rng.uniform(size=P).astype(np.float32)
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda")
    ens.q.copy_(torch.tensor(q0))
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=pe)
    assert hmc.integrator.numSteps == L
    acc = torch.empty(P, dtype=torch.uint8, device="cuda")
    pout = torch.empty_like(ens.q)
    hmc.step(1 / KB, accept=acc, z=torch.tensor(z, device="cuda"), u=torch.tensor(u, device="cuda"), p_out=pout)
    torch.cuda.synchronize()
    sel = np.sort(rng.choice(P, 128, replace=False))
    po = O.Logistic(X, y, 1.0)
    qr, pr, accr, oh, nh = O.hmc_iter(q0[:, sel].astype(np.float64), z[:, sel].astype(np.float64),
                                      u[sel].astype(np.float64), np.ones(128), 1 / KB, h, L, po, bug_compat=False)
    a = acc.cpu().numpy().astype(bool)[sel]
    with np.errstate(over="ignore"):
        clear = np.abs(u[sel] - np.minimum(1, np.exp(oh - nh))) > TIE[np.float32]
    assert np.array_equal(a[clear], accr[clear])
    assert 5 < a.sum() < 128  # both outcomes occur
    same = a == accr
    qg = ens.q.cpu().numpy()[:, sel]
    print("config 3 full size: rel err q", rel_err(qg[:, same], qr[:, same]), "accepted", int(a.sum()), "of 128")
    assert rel_err(qg[:, same], qr[:, same]) < 1e-5
    acc_same = same & a
    assert rel_err(pout.cpu().numpy()[:, sel][:, acc_same], pr[:, acc_same]) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16x3"])
def test_logistic_endpoint_cache_is_exact(E, precision):
    """EHMC_FLAG_REUSE_ENDPOINT: starting every trajectory from the gradient / energy kept at the end of the
    previous one (accepted: the trajectory's end, rejected: its start) gives bit-identical chains with one
    gradient launch less per iteration."""
    import torch

    rng = np.random.RandomState(41)
    N, D, P, L = 600, 40, 700, 4
    X = rng.standard_normal((N, D)) / np.sqrt(D)
    y = (rng.uniform(size=N) < 0.5).astype(np.float64)
    pot = E.LogisticPotential(X, y, 1.0, precision=precision)
    q0 = torch.tensor(rng.standard_normal((D, P)), dtype=torch.float32, device="cuda")
    ctx = E._lib.Context.get()
    res = {}
    for reuse in (False, True):
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=3)
        ens.q.copy_(q0)
        hmc = E.HMC(ens, L * 0.3 + 1e-9, 0.3, None, potential=pot, seed=3, bugCompat=False)
        acc = torch.empty(P, dtype=torch.uint8, device="cuda")
        n0 = ctx.launch_count()
        tot = 0
        for it in range(6):
            hmc.step(1 / KB, accept=acc, reuseEndpoint=reuse)
            tot += int(acc.sum().item())
            if it == 2:
                # an external change of q invalidates the promise: the caller must drop the flag once
                ens.q.mul_(1.0)
        res[reuse] = (ens.q.clone(), tot, ctx.launch_count() - n0)
    assert torch.equal(res[False][0], res[True][0])
    assert res[False][1] == res[True][1] and 0.2 * 6 * P < res[True][1] < 6 * P  # accepts and rejects both occur
    assert res[False][2] - res[True][2] == 5  # one gradient launch less in every iteration but the first


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_nbody_endpoint_cache_is_exact(E, dt):
    """The same promise for the pairwise-gravity family: the first all-pairs sweep of every iteration is
    replaced by the forces / energy kept from the previous one; chains stay bit-identical."""
    import torch

    rng = np.random.RandomState(43)
    B, P, L = 24, 150, 3
    pot = E.NBodyPotential(rng.uniform(0.5, 1.5, B) / B, G=1.0, eps=0.1)
    tdt = torch.float32 if dt == np.float32 else torch.float64
    q0 = torch.tensor(rng.standard_normal((3 * B, P)), dtype=tdt, device="cuda")
    res = {}
    for reuse in (False, True):
        ens = E.Ensemble(3 * B, P, dtype=dt, device="cuda", seed=3)
        ens.q.copy_(q0)
        hmc = E.HMC(ens, L * 0.2 + 1e-9, 0.2, None, potential=pot, seed=3, bugCompat=False)
        acc = torch.empty(P, dtype=torch.uint8, device="cuda")
        st = torch.zeros(2 * 3 * B + 3, dtype=torch.float64, device="cuda")
        tot, hsum = 0, 0.0
        for it in range(7):
            hmc.step(1 / KB, accept=acc, stats=st, reuseEndpoint=reuse)
            tot += int(acc.sum().item())
            hsum += float(st[2].item())
        res[reuse] = (ens.q.clone(), tot, hsum)
    assert torch.equal(res[False][0], res[True][0])
    assert res[False][1] == res[True][1] and 0.1 * 7 * P < res[True][1] < 7 * P
    assert res[False][2] == res[True][2]  # the Hamiltonians entering the statistics are the same numbers


def test_logistic_tensor_core_hmc_iteration(E):
    """A whole HMC iteration driven by the tensor-core gradient stays close to the exact one."""
    import torch

    N, D, P, L, h = 2048, 256, 256, 10, 0.02
    rng = np.random.RandomState(9)
    X = rng.standard_normal((N, D)) / np.sqrt(D)
    y = (rng.uniform(size=N) < 0.5).astype(np.float64)
    q0 = rng.standard_normal((D, P))
    z = rng.standard_normal((D, P))
    u = rng.uniform(size=P)
    qr, pr, accr, oh, nh = O.hmc_iter(q0, z, u, np.ones(P), 1 / KB, h, L, O.Logistic(X, y, 1.0))
    ens = E.Ensemble(D, P, dtype=np.float32, device="cuda")
    ens.q.copy_(torch.tensor(q0, dtype=torch.float32))
    hmc = E.HMC(ens, L * h + 1e-9, h, None, potential=E.LogisticPotential(X, y, 1.0, precision="bf16"))
    acc = torch.empty(P, dtype=torch.uint8, device="cuda")
    hmc.step(1 / KB, accept=acc, z=torch.tensor(z, dtype=torch.float32, device="cuda"),
             u=torch.tensor(u, dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    a = acc.cpu().numpy().astype(bool)
    same = a == accr
    assert same.mean() > 0.9
    assert rel_err(ens.q.cpu().numpy()[:, same], qr[:, same]) < 2e-3


def test_checkpoint_resume_is_bit_exact(E, tmp_path):
    """Philox is counter based: a run resumed from a checkpoint equals the uninterrupted run."""
    import torch

    D, P = 10, 4096

    def make():
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=3)
        ens.setPosition(1.0)
        return ens, E.HMC(ens, 0.4, 0.05, None, potential=E.FunnelPotential(D, 3.0), seed=3)

    ens_a, hmc_a = make()
    hmc_a.run(6, 1 / KB, collectStats=False)
    E.io.saveCheckpoint(str(tmp_path / "ck"), hmc_a)
    hmc_a.run(5, 1 / KB, collectStats=False)
    ens_b, hmc_b = make()
    ens_b.q.zero_()
    E.io.loadCheckpoint(str(tmp_path / "ck"), hmc_b)
    hmc_b.run(5, 1 / KB, collectStats=False)
    torch.cuda.synchronize()
    assert torch.equal(ens_a.q, ens_b.q)


def test_checkpoint_resume_of_an_adaptive_run(E, tmp_path):
    """A checkpoint taken between two run() calls of an ADAPTIVE driver (step size adapted at fixed numSteps, so
    simulTime = numSteps * stepSize has drifted from the constructor's value; mass scales set) restores the trajectory
    length, the flags and the scales: the resumed driver continues exactly like the uninterrupted one."""
    import torch

    D, P = 10, 4096

    def make(bugCompat=False):
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=5)
        ens.setPosition(1.0)
        return ens, E.HMC(ens, 0.81, 0.05, None, potential=E.HarmonicPotential(np.logspace(-1, 1, D)), seed=5,
                          bugCompat=bugCompat, rejectNonFinite=True)

    ens_a, hmc_a = make()
    hmc_a.run(60, 1 / KB, adapt=True, adaptMass=True, keepNumSteps=True)
    assert hmc_a.massScale is not None and hmc_a.integrator.numSteps == 16
    assert abs(hmc_a.simulTime - 0.81) > 1e-3  # drifted with the step size
    E.io.saveCheckpoint(str(tmp_path / "ck"), hmc_a)
    ra = hmc_a.run(5, 1 / KB)
    ens_b, hmc_b = make(bugCompat=True)  # flags differ on purpose: the checkpoint's win
    hmc_b.rejectNonFinite = None
    E.io.loadCheckpoint(str(tmp_path / "ck"), hmc_b)
    assert hmc_b.integrator.numSteps == 16 and hmc_b.simulTime == hmc_a.integrator.numSteps * hmc_b.stepSize
    assert hmc_b.bugCompat is False and hmc_b.rejectNonFinite is True
    np.testing.assert_array_equal(hmc_b.massScale, hmc_a.massScale)
    rb = hmc_b.run(5, 1 / KB)
    torch.cuda.synchronize()
    assert torch.equal(ens_a.q, ens_b.q)
    assert ra["acceptRate"] == rb["acceptRate"]


def test_set_position_draws_fresh_positions_each_call(E):
    """Device-backed Ensemble.setPosition draws NEW positions on every call, like the reference's norm.rvs
    (src/ensemble.py:72-74); the first call is the documented stream of (seed, iteration word ~0)."""
    import torch

    ens = E.Ensemble(3, 1000, dtype=np.float32, device="cuda", seed=9)
    q1 = ens.setPosition(2.0).clone()
    q2 = ens.setPosition(2.0).clone()
    assert not torch.equal(q1, q2)
    assert abs(float(q2.std()) - 2.0) < 0.1 and abs(float(torch.corrcoef(torch.stack([q1.flatten(), q2.flatten()]))[0, 1])) < 0.1
    ref = torch.empty_like(q1)
    E._lib.set_position(E._lib.Context.get(), ref, 2.0, 9, 0)
    assert torch.equal(q1, ref)


def test_run_rejects_non_finite_ratios_by_default(E):
    """HMC.run must survive divergent trajectories: particles deep in the funnel's neck (v = -60: e^{60} x^2 overflows,
    the energy difference is inf - inf = NaN) are REJECTED by run() and stay where they were, while step() keeps the
    reference's rule -- u > NaN is False, the NaN proposal is accepted (src/HMC.py:168-173)."""
    import torch

    D, P = 10, 2048
    q0 = torch.randn((D, P), dtype=torch.float32, device="cuda")
    q0[0, :64] = -60.0
    q0[1:, :64] = 1.0e10

    def make(**kw):
        ens = E.Ensemble(D, P, dtype=np.float32, device="cuda", seed=2)
        ens.q.copy_(q0)
        return ens, E.HMC(ens, 0.5, 0.05, None, potential=E.FunnelPotential(D, 3.0), seed=2, bugCompat=False, **kw)

    ens, hmc = make()
    r = hmc.run(3, 1 / KB)
    torch.cuda.synchronize()
    assert torch.isfinite(ens.q).all()
    assert torch.equal(ens.q[:, :64], q0[:, :64])  # never moved
    assert np.isfinite(r["mean"].numpy()).all()
    ens2, hmc2 = make()  # reference semantics outside run()
    hmc2.step(1 / KB)
    torch.cuda.synchronize()
    assert not torch.isfinite(ens2.q[:, :64]).all()
    ens3, hmc3 = make(rejectNonFinite=False)  # explicit opt-out applies to run() too
    hmc3.run(1, 1 / KB)
    torch.cuda.synchronize()
    assert not torch.isfinite(ens3.q[:, :64]).all()


def test_tensor_of_another_device_is_refused(E):
    """The C-ABI checks DLTensor.device.device_id against the context's device (and the device tensors of one call
    against each other): a tensor that claims to live on another GPU is refused with EHMC_ERR_INVALID instead of
    being dereferenced on the wrong device, and the refusal leaves no state behind."""
    import ctypes

    import torch

    q = torch.ones((2, 8), dtype=torch.float64, device="cuda")
    e = torch.empty(8, dtype=torch.float64, device="cuda")
    pot = E.HarmonicPotential([1.0, 2.0])
    ctx = E._lib.Context.get()
    view = E._lib.DL(q)
    dev_id = ctypes.c_int32.from_address(view.ptr + 12)  # DLTensor: void* data; int32 device_type; int32 device_id
    assert dev_id.value == q.device.index
    dev_id.value = q.device.index + 1
    try:
        with pytest.raises(E._lib.EhmcError) as ei:
            E._lib.potential_eval(ctx, pot.handle(64, ctx), view, e, None)
        assert "device" in str(ei.value).lower()
        with pytest.raises(E._lib.EhmcError):
            E._lib.potential_eval(ctx, pot.handle(64, ctx), view, None, None)  # alone: wrong device for the context
    finally:
        dev_id.value = q.device.index
    E._lib.potential_eval(ctx, pot.handle(64, ctx), view, e, None)
    torch.cuda.synchronize()
    assert torch.allclose(e, torch.full_like(e, 1.5))
