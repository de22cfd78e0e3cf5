"""Generates tests/golden/*.npz by executing the UNMODIFIED reference
(/root/reference/src, via oracle/ref_runner.py and the jax stand-in).

Run in the build container only:  python tests/golden/make_golden.py
The fixtures are committed; the GPU box never needs /root/reference.
All cases are seeded with NumPy's legacy global MT19937 (the reference's RNG).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import hmc_oracle as O  # noqa: E402  closed-form potentials for the unpinned models
from oracle import ref_runner as R  # noqa: E402

KB = 1.380649e-23
SEED = 20221018


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print("wrote", name, {k: np.shape(v) for k, v in arrs.items()})


def known_answers():
    ensemble, integrator, potential, HMC = R.modules()
    # KA1 src/tests/test_potential.py:14-25
    ens = ensemble.Ensemble(2, 10)
    ens.q[:, 0] = np.array([3.0, 4.0])
    ka1 = potential.harmonicPotentialND(ens.q, np.array([2, 3]))
    # KA2 src/tests/test_ensemble.py:26-44
    e2 = ensemble.Ensemble(4, 100)
    q, p, m, w = e2.particle(10)
    try:
        e2.particle(101)
        idx_err = 0
    except IndexError:
        idx_err = 1
    # row N: getAccelNBody on a random (3, 7) system, every i
    rng = np.random.RandomState(SEED)
    qn = rng.standard_normal((3, 7))
    mn = rng.uniform(0.5, 2.0, 7)
    acc = np.stack([potential.getAccelNBody(qn, mn, i) for i in range(7)], axis=1)
    # row N1: gravitationalPotential sign (reference returns +G m1 m2 / r)
    gp = potential.gravitationalPotential(qn[:, 0], qn[:, 1], mn[0], mn[1])
    nbp = potential.nBodyPotential(qn, mn)
    # int(finalTime/stepSize) cases (row F)
    ns = np.array([int(0.3 / 0.1), int(1.0 / 0.05), int(2.5 / 0.05), int(0.5 / 0.05)])
    save("known_answers", ka1=ka1, ka2_q=q, ka2_p=p, ka2_m=m, ka2_w=w, ka2_index_error=idx_err,
         nbody_q=qn, nbody_m=mn, nbody_acc=acc, grav_pot_01=gp, nbody_pot=nbp, num_steps=ns)


def integrators():
    k = np.array((2.0, 3.0))
    grad = lambda q: k * q  # analytic gradient of harmonicPotentialND (potential.py:27)
    rng = np.random.RandomState(SEED + 1)
    D, P = 2, 16
    final_time = 2 * np.pi / np.sqrt(2.0)  # tests/test_integrator_harmonic.py:51-55
    for tag, mass in (("unitmass", np.ones(P)), ("randmass", rng.uniform(0.5, 2.0, P))):
        q0 = rng.standard_normal((D, P)) * 10.0
        p0 = rng.standard_normal((D, P)) * np.sqrt(mass * KB * 1000.0 / KB)
        for method in ("Leapfrog", "Stormer-Verlet"):
            for h in (0.1, 0.01):
                (out,), L = R.run_integrate(q0, p0, mass, h, final_time, grad, method)
                save(f"integrate_{method.replace('-', '')}_{tag}_h{h}", q0=q0, p0=p0, mass=mass, k=k,
                     h=h, final_time=final_time, L=L, q1=out[0], p1=out[1])
    # 3-D, the benchmark script's setup (tests/test_integrator_benchmarks_harmonic.py:25-37), smaller P
    k3 = np.array((2.0, 3.0, 4.0))
    q0 = rng.standard_normal((3, 12)) * 5.0
    p0 = rng.standard_normal((3, 12))
    (out,), L = R.run_integrate(q0, p0, np.ones(12), 0.1, 10, lambda q: k3 * q, "Leapfrog")
    save("integrate_Leapfrog_3d_bench", q0=q0, p0=p0, mass=np.ones(12), k=k3, h=0.1, final_time=10.0,
         L=L, q1=out[0], p1=out[1])


def nbody_reference_mode():
    # tests/test_integrator_solar_system.py:28-46 (gradient=None -> bodies are the particles)
    mass = np.array([5.972e24, 1.989e30, 7.34e22])
    q0 = np.zeros((3, 3))
    p0 = np.zeros((3, 3))
    q0[:, 0] = [1.52e11, 0, 0]
    p0[:, 0] = np.array([0, 29800, 0]) * mass[0]
    q0[:, 2] = [1.52e11, 3.844e8, 0]
    p0[:, 2] = np.array([0, 29800, 1022]) * mass[2]
    for method in ("Leapfrog", "Stormer-Verlet"):
        outs, L = R.run_integrate(q0, p0, mass, 600, 3600, None, method, calls=3)
        save(f"nbody_mode_{method.replace('-', '')}", q0=q0, p0=p0, mass=mass, h=600.0, final_time=3600.0,
             L=L, q=np.stack([o[0] for o in outs]), p=np.stack([o[1] for o in outs]))


def hmc_case(name, D, P, S, pot, temperature, q_std, simul_time, h, mass=None, method="Leapfrog",
             potential=None, extra=None):
    potential = potential or (lambda q: pot.energy(q))
    out = R.run_get_samples(SEED, D, P, mass, potential, pot.grad, S, temperature, q_std, simul_time, h,
                            method)
    save(name, D=D, P=P, S=S, temperature=temperature, q_std=q_std, simul_time=simul_time, h=h,
         L=out["numSteps"], mass=out["mass"], z_init=out["z_init"], z=out["z"], u=out["u"],
         samples=out["samples"], momenta=out["momenta"], **(extra or {}))


def hmc_cases():
    _, _, potential, _ = R.modules()
    rng = np.random.RandomState(SEED + 2)
    # C1 shape reduced: the reference's OWN potential function, k = (1, 1)
    k1 = np.array((1.0, 1.0))
    hmc_case("hmc_iso2d", 2, 64, 12, O.DiagGaussian(k1), 1 / KB, 1.0, 1.0, 0.05,
             potential=lambda q: potential.harmonicPotentialND(q, k1), extra=dict(k=k1))
    hmc_case("hmc_iso2d_rough", 2, 64, 10, O.DiagGaussian(k1), 1 / KB, 2.0, 3.6, 0.9,
             potential=lambda q: potential.harmonicPotentialND(q, k1), extra=dict(k=k1))
    k2 = np.array((2.0, 3.0))
    hmc_case("hmc_diag2d_randmass", 2, 40, 8, O.DiagGaussian(k2), 1 / KB, 2.0, 0.6, 0.1,
             mass=rng.uniform(0.5, 2.0, 40), potential=lambda q: potential.harmonicPotentialND(q, k2),
             extra=dict(k=k2))
    hmc_case("hmc_iso2d_stormer", 2, 32, 6, O.DiagGaussian(k1), 1 / KB, 1.0, 1.0, 0.05,
             method="Stormer-Verlet", potential=lambda q: potential.harmonicPotentialND(q, k1),
             extra=dict(k=k1))
    # tests/test_HMC.py:110-130 (test2): mean (5,5), cov [[4,-3],[-3,4]], T = 300 K and T = 1/kB
    from scipy.stats import multivariate_normal

    mean = np.ones(2) * 5
    cov = np.array([[4.0, -3.0], [-3.0, 4.0]])
    prec = np.linalg.inv(cov)
    pot = O.DenseGaussian(prec, mean)
    logpdf = lambda q: -multivariate_normal.logpdf(q, mean, cov=cov)
    hmc_case("hmc_corr2d_T300", 2, 50, 4, pot, 300, 1, 0.5, 0.05, potential=logpdf,
             extra=dict(prec=prec, mean=mean))
    hmc_case("hmc_corr2d_Tinv", 2, 50, 10, pot, 1 / KB, 1, 0.5, 0.05, potential=logpdf,
             extra=dict(prec=prec, mean=mean))
    # C2 shape reduced: 100-D dense precision, L = 50 (parity unpinned model, reference loop)
    A = rng.standard_normal((100, 100))
    prec100 = A @ A.T / 100 + np.eye(100)
    hmc_case("hmc_dense100", 100, 6, 3, O.DenseGaussian(prec100), 1 / KB, 1.0, 50 * 0.05, 0.05,
             extra=dict(prec=prec100, mean=np.zeros(100)))
    hmc_case("hmc_dense100_rough", 100, 12, 3, O.DenseGaussian(prec100), 1 / KB, 1.0, 8 * 0.55, 0.55,
             extra=dict(prec=prec100, mean=np.zeros(100)))
    mean20 = rng.standard_normal(20)
    A = rng.standard_normal((20, 20))
    prec20 = A @ A.T / 20 + np.eye(20)
    hmc_case("hmc_dense20_mean", 20, 10, 4, O.DenseGaussian(prec20, mean20), 1 / KB, 1.0, 1.0, 0.1,
             mass=rng.uniform(0.5, 2.0, 10), extra=dict(prec=prec20, mean=mean20))
    # C5 shape reduced: Neal's funnel, 10-D
    hmc_case("hmc_funnel10", 10, 32, 5, O.Funnel(10, 3.0), 1 / KB, 1.0, 0.4, 0.02,
             extra=dict(sigma_v=3.0))
    hmc_case("hmc_funnel10_rough", 10, 48, 5, O.Funnel(10, 3.0), 1 / KB, 1.0, 1.5, 0.25,
             extra=dict(sigma_v=3.0))
    # C4 shape reduced: 6 bodies x 3-D per particle, G = 1, m = 1/B, eps = 0 and eps = 0.05
    B = 6
    for eps in (0.0, 0.05):
        nb = O.NBody(np.ones(B) / B, G=1.0, eps=eps)
        hmc_case(f"hmc_nbody6_eps{eps}", 3 * B, 4, 3, nb, 1 / KB, 1.0, 0.1, 0.01,
                 extra=dict(body_mass=nb.m, G=1.0, eps=eps))
    # C3 shape reduced: logistic regression
    N, Dl = 40, 6
    X = rng.standard_normal((N, Dl)) / np.sqrt(Dl)
    theta = rng.standard_normal(Dl)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-X @ theta))).astype(np.float64)
    hmc_case("hmc_logistic", Dl, 8, 3, O.Logistic(X, y, 1.0), 1 / KB, 1.0, 0.5, 0.05,
             extra=dict(X=X, y=y, prior_scale=1.0))
    hmc_case("hmc_logistic_rough", Dl, 16, 4, O.Logistic(X, y, 1.0), 1 / KB, 1.0, 2.4, 0.8,
             extra=dict(X=X, y=y, prior_scale=1.0))


if __name__ == "__main__":
    assert R.available(), "run in the build container (needs /root/reference)"
    known_answers()
    integrators()
    nbody_reference_mode()
    hmc_cases()
