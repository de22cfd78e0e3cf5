"""Runs the UNMODIFIED reference (``/root/reference/src``) -- TEST INFRASTRUCTURE ONLY.

Works only in the build container (the GPU box has no ``/root/reference``); it is
used by ``tests/golden/make_golden.py`` to produce the committed fixtures and by
the optional ``tests/test_oracle_vs_reference.py`` (skipped when the reference
tree is absent).  JAX is replaced by the NumPy alias in ``oracle/jax_standin``
(SURVEY.md section 8c); every gradient is passed explicitly so ``jax.grad`` is
never on the pinned path (src/HMC.py:57-58 allows that).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

REFERENCE_SRC = os.environ.get("EHMC_REFERENCE_SRC", "/root/reference/src")
_STANDIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jax_standin")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "integrator.py"))


def modules():
    """Import the reference's flat modules (ensemble, integrator, potential, HMC)."""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    for p in (_STANDIN, REFERENCE_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ensemble  # noqa: E402
    import integrator  # noqa: E402
    import potential  # noqa: E402
    import HMC  # noqa: E402

    return ensemble, integrator, potential, HMC


class _RecordingNorm:
    """Wraps scipy.stats.norm inside the reference's ``ensemble`` module and records
    the STANDARD normals behind every ``norm.rvs(scale=, size=)`` call by replaying
    NumPy's global MT19937 state; asserts bitwise that rvs == standard_normal*scale
    (SURVEY row B/C)."""

    def __init__(self, real):
        self._real = real
        self.draws = []

    def rvs(self, scale=1.0, size=None):
        state = np.random.get_state()
        out = self._real.rvs(scale=scale, size=size)
        after = np.random.get_state()
        np.random.set_state(state)
        z = np.random.standard_normal(size)
        assert np.array_equal(z * scale, out), "norm.rvs != standard_normal*scale"
        np.random.set_state(after)
        self.draws.append(z)
        return out


@contextlib.contextmanager
def recording():
    """Context manager: records z draws (ensemble.norm.rvs) and u draws
    (np.random.uniform) made by the reference while active."""
    ensemble, _, _, _ = modules()
    rec = _RecordingNorm(ensemble.norm)
    real_norm = ensemble.norm
    real_uniform = np.random.uniform
    uniforms = []

    def uniform(*a, **k):
        out = real_uniform(*a, **k)
        uniforms.append(np.array(out, copy=True))
        return out

    ensemble.norm = rec
    np.random.uniform = uniform
    try:
        yield rec.draws, uniforms
    finally:
        ensemble.norm = real_norm
        np.random.uniform = real_uniform


def run_get_samples(seed, D, P, mass, potential, gradient, num_samples, temperature, q_std,
                    simul_time, step_size, method="Leapfrog"):
    """HMC(...).getSamples(...) of the reference with a seeded global stream.
    Returns dict(samples, momenta, q_init_z, z[list], u[list], numSteps)."""
    ensemble, _, _, HMC = modules()
    np.random.seed(seed)
    ens = ensemble.Ensemble(D, P)
    if mass is not None:
        ens.mass = np.asarray(mass, dtype=np.float64)
    with recording() as (zs, us), contextlib.redirect_stdout(io.StringIO()):
        h = HMC.HMC(ens, simul_time, step_size, None, potential=potential, gradient=gradient,
                    method=method)
        samples, momenta = h.getSamples(num_samples, temperature, q_std)
    return dict(samples=samples, momenta=momenta, z_init=zs[0], z=np.stack(zs[1:]),
                u=np.stack(us), numSteps=h.integrator.numSteps, mass=ens.mass.copy())


def run_integrate(q0, p0, mass, step_size, final_time, gradient, method="Leapfrog", calls=1):
    """Leapfrog / StormerVerlet .integrate() of the reference on given state."""
    ensemble, integrator, _, _ = modules()
    D, P = q0.shape
    ens = ensemble.Ensemble(D, P)
    ens.mass = np.asarray(mass, dtype=np.float64).copy()
    ens.q[:] = q0
    ens.p[:] = p0
    cls = integrator.Leapfrog if method == "Leapfrog" else integrator.StormerVerlet
    with contextlib.redirect_stdout(io.StringIO()):
        integ = cls(ens, step_size, final_time, gradient)
    out = []
    for _ in range(calls):
        q, p = integ.integrate()
        assert q is ens.q and p is ens.p  # in-place aliasing
        out.append((q.copy(), p.copy()))
    return out, integ.numSteps
