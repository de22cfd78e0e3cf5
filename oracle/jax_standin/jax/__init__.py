"""Minimal stand-in for the ``jax`` package -- TEST INFRASTRUCTURE ONLY.

JAX / NumPyro are not installable in the build image (no network, no wheel).
The reference (``/root/reference/src/*.py``) touches only nine ``jax.numpy``
names plus ``jax.grad`` / ``jax.config.update`` / ``jit`` / ``pmap``
(SURVEY.md section 8c).  With x64 enabled those are plain IEEE-double NumPy
operations, so aliasing them to NumPy lets the UNMODIFIED reference files
execute in float64.  This package is put on ``sys.path`` only by
``oracle/ref_runner.py`` (golden-vector generation in the build container);
the product never imports it.
"""
import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import scipy  # noqa: F401  (jax.scipy)


class _Config:
    def update(self, *_a, **_k):  # jax.config.update("jax_enable_x64", True)
        return None


config = _Config()


def jit(f=None, **_k):
    return f if f is not None else (lambda g: g)


def pmap(f=None, **_k):
    return f if f is not None else (lambda g: g)


def grad(f, h=1e-6):
    """Numerical stand-in for jax.grad (4th-order central difference).

    Only used for ad-hoc cross-checks; every golden vector is generated with an
    explicit analytic ``gradient=`` so that this function is never on the
    pinned path (the reference allows that: HMC.py:57-58).
    """

    def g(x):
        x = _np.asarray(x, dtype=_np.float64)
        out = _np.zeros_like(x)
        for i in range(x.size):
            e = _np.zeros_like(x)
            e.flat[i] = h
            out.flat[i] = (
                -f(x + 2 * e) + 8 * f(x + e) - 8 * f(x - e) + f(x - 2 * e)
            ) / (12 * h)
        return out

    return g
