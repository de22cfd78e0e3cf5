"""Data formats either side of the hot path ("next" rows 3 and 4).

* N-body input files of the reference's sample lab (samples/NBody/pl2.txt, pl3.txt, pl100.txt,
  pl1k.txt; reader samples/NBody/MiscFunctions.py:8-43): header "N tmax dt", then N masses, N
  positions, N velocities.
* A resumable checkpoint of an HMC run: positions, masses, Philox (seed, iteration), step size,
  trajectory length (simulTime, numSteps), flags and mass scales.  Because the RNG is counter based, a
  driver restored between two run() / step() / getSamples() calls continues exactly like the one that
  was never interrupted (bit for bit, adaptive runs included: the Robbins-Monro gain restarts with every
  run() call, interrupted or not, so there is no adapter state to carry).
"""
from __future__ import annotations

import numpy as np


def readNBodyInput(fname):
    """Returns [N, tmax, dt, M (N,), PScoord (N, 2, 3)] like ReadInput of the reference's
    samples/NBody/MiscFunctions.py:8-43 (PScoord[j, 0] = position, PScoord[j, 1] = velocity)."""
    with open(fname, "r") as f:
        tokens = f.read().split()
    N = int(tokens[0])
    tmax = float(tokens[1])
    dt = float(tokens[2])
    vals = np.array(tokens[3:3 + 7 * N], dtype=np.float64)
    if vals.size != 7 * N:
        raise ValueError(f"{fname}: expected {7 * N} numbers after the header, found {vals.size}")
    M = vals[:N].copy()
    PScoord = np.zeros((N, 2, 3))
    PScoord[:, 0, :] = vals[N:4 * N].reshape(N, 3)
    PScoord[:, 1, :] = vals[4 * N:7 * N].reshape(N, 3)
    return [N, tmax, dt, M, PScoord]


def nbodyInputToEnsembleColumn(M, PScoord):
    """One ensemble particle (column) of the NBodyPotential family from a lab input: positions
    flattened component-major d = c*N + b (src/potential.py:83-84), momenta = m_b * v_b."""
    q = np.ascontiguousarray(PScoord[:, 0, :].T).reshape(-1)
    p = np.ascontiguousarray((PScoord[:, 1, :] * M[:, None]).T).reshape(-1)
    return q, p


def saveCheckpoint(path, hmc):
    """Resumable state of an HMC driver (host or device ensemble)."""
    ens = hmc.ensemble
    to_np = (lambda a: a) if not ens.onDevice else (lambda a: a.detach().cpu().numpy())
    np.savez_compressed(path, q=to_np(hmc.integrator.q), mass=to_np(ens.mass), seed=np.uint64(hmc.seed),
                        iteration=np.uint64(hmc.iteration), stepSize=np.float64(hmc.stepSize),
                        simulTime=np.float64(hmc.simulTime), numSteps=np.int64(hmc.integrator.numSteps),
                        particleOffset=np.int64(ens.particleOffset), method=np.array(hmc.method),
                        bugCompat=np.int8(hmc.bugCompat),
                        rejectNonFinite=np.int8(-1 if hmc.rejectNonFinite is None else int(hmc.rejectNonFinite)),
                        massScale=np.zeros(0) if getattr(hmc, "massScale", None) is None else np.asarray(hmc.massScale))


def loadCheckpoint(path, hmc):
    """Restores positions, masses, Philox counter and step size into a compatible HMC driver."""
    with np.load(path if str(path).endswith(".npz") else str(path) + ".npz") as f:
        q, mass = f["q"], f["mass"]
        ens = hmc.ensemble
        if q.shape != (ens.numDimensions, ens.numParticles):
            raise ValueError(f"checkpoint holds {q.shape}, the ensemble is {(ens.numDimensions, ens.numParticles)}")
        if ens.onDevice:
            import torch

            ens.q.copy_(torch.from_numpy(q).to(ens.q.dtype))
            ens.mass.copy_(torch.from_numpy(mass).to(ens.mass.dtype))
        else:
            ens.q[...] = q
            ens.mass[...] = mass
        hmc.integrator.q = ens.q
        hmc.integrator.mass = ens.mass
        hmc.seed = int(f["seed"])
        hmc.iteration = int(f["iteration"])
        hmc.stepSize = float(f["stepSize"])
        hmc.integrator.stepSize = hmc.stepSize
        if "numSteps" in f:
            # the trajectory length the saved driver had (run(keepNumSteps=True) lets simulTime = numSteps * stepSize
            # drift away from the constructor's value)
            hmc.simulTime = float(f["simulTime"])
            hmc.integrator.finalTime = hmc.simulTime
            hmc.integrator.numSteps = int(f["numSteps"])
            hmc.bugCompat = bool(f["bugCompat"])
            rnf = int(f["rejectNonFinite"])
            hmc.rejectNonFinite = None if rnf < 0 else bool(rnf)
            ms = f["massScale"]
            hmc.massScale = None if ms.size == 0 else np.array(ms, dtype=np.float64)
        else:  # checkpoints written before these fields existed
            hmc.integrator.numSteps = int(hmc.simulTime / hmc.stepSize)
    return hmc
