"""Integrators -- host mirror of the reference's ``src/integrator.py``.

``Leapfrog.integrate()`` / ``StormerVerlet.integrate()`` keep the reference's
signature and in-place semantics (they mutate ``ensemble.q`` / ``ensemble.p`` and
return the same objects, src/integrator.py:94-123,126-165) but run the whole
trajectory -- every particle, all ``numSteps`` steps -- as ONE fused CUDA kernel
through ``ehmc_leapfrog`` / ``ehmc_stormer_verlet``.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .ensemble import Ensemble  # noqa: F401  (re-exported like the reference module does)
from .potential import _descriptor, gravConst

G = gravConst  # "from scipy.constants import G  # for debug" (src/integrator.py:16)


class Integrator:
    """Positions and momenta at final simulation time ``finalTime`` for a system of
    N particles moving in a given potential (src/integrator.py:20-59)."""

    _method = None

    def __init__(self, ensemble, stepSize, finalTime, gradient):
        self.ensemble = ensemble
        # initial positions / momenta: the SAME array objects as the ensemble's (:40-43)
        self.q = ensemble.q
        self.p = ensemble.p
        self.mass = ensemble.mass
        self.numParticles = ensemble.numParticles
        self.stepSize = stepSize
        self.finalTime = finalTime
        # float floor exactly like the reference (src/integrator.py:51); never recomputed in C
        self.numSteps = int(self.finalTime / self.stepSize)
        self.gradient = gradient
        self.nBodyMode = not gradient
        if self.nBodyMode:
            # src/integrator.py:57-59: particles are the bodies, acceleration = getAccelNBody
            print(f"Gradient={gradient} - performing nBody simulation.")
            self.potential = None
        else:
            self.potential = _descriptor(gradient)
            if self.potential.numDimensions != ensemble.numDimensions:
                raise ValueError(
                    f"potential has {self.potential.numDimensions} dimensions, ensemble has {ensemble.numDimensions}")

    @property
    def v(self):
        """Velocities p / mass (src/integrator.py:45,106,120)."""
        return self.p / self.mass

    def getAccel(self, i):
        """Acceleration of the i-th particle, -gradient(q[:, i]) / mass[i] (src/integrator.py:61-73)."""
        if self.nBodyMode:
            return self.getAccelNBody(i)
        return -self.potential.gradient(self.q[:, i]) / self.mass[i]

    def getAccelNBody(self, i):
        """src/integrator.py:75-85."""
        from .potential import getAccelNBody

        return getAccelNBody(self.q, self.mass, i)

    def integrate(self):
        raise NotImplementedError("Integrator superclass doesn't specify integration method")

    # ------------------------------------------------------------------------------------
    def _arrays(self):
        q, p, m = self.q, self.p, self.mass
        if isinstance(q, np.ndarray):
            dt = q.dtype
            if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
                raise TypeError("q must be float32 or float64")
            if not isinstance(p, np.ndarray) or p.dtype != dt or p.shape != q.shape:
                raise TypeError("p must be a NumPy array with q's shape and dtype")
            m = np.ascontiguousarray(m, dtype=dt)
            if m.shape != (q.shape[1],):
                raise ValueError("mass must have one entry per particle")
            return q, p, m, dt.itemsize * 8
        if p.dtype != q.dtype or p.shape != q.shape or m.dtype != q.dtype:
            raise TypeError("q, p and mass must be CUDA tensors of one dtype and matching shapes")
        return q, p, m, q.element_size() * 8

    def _run(self, stormer):
        q, p, m, bits = self._arrays()
        ctx = _lib.Context.get(q.device.index if not isinstance(q, np.ndarray) else None)
        stream = _lib.current_stream_ptr(q)
        if self.nBodyMode:
            _lib.integrate_nbody_mode(ctx, _lib.STORMER_VERLET if stormer else _lib.LEAPFROG, q, p, m, gravConst,
                                      self.stepSize, self.stepSize**2, self.numSteps, stream)
        else:
            _lib.leapfrog(ctx, self.potential.handle(bits, ctx), q, p, m, self.stepSize, self.stepSize**2,
                          self.numSteps, stream, stormer=stormer)
            if isinstance(q, np.ndarray) and bits == 32:
                # host-backed calls are synchronous anyway: surface a divergent trajectory that left the fp16 operand
                # range of the tensor-core dense kernel (device-backed callers poll Context.overflow_count themselves)
                n = ctx.overflow_count()
                if n:
                    raise FloatingPointError(
                        f"{n} particle trajectories diverged past the operand range of the float32 tensor-core "
                        "kernel (step size above the stability limit?); their q, p are not valid. Re-run with "
                        "Context.set_option('dense_path', 1) for the exact CUDA-core kernel")
        # positions and momenta of all particles at finalTime: the same objects, mutated in place
        return (self.q, self.p)


class Leapfrog(Integrator):
    _method = "Leapfrog"

    def integrate(self):
        """Velocity-Verlet ("leap frog") in position/acceleration form,
        src/integrator.py:94-123; numSteps + 1 gradient evaluations per particle."""
        return self._run(stormer=False)


class StormerVerlet(Integrator):
    _method = "Stormer-Verlet"

    def integrate(self):
        """Two-step position Verlet, src/integrator.py:126-165."""
        return self._run(stormer=True)
