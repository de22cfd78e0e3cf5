"""Multi-GPU plumbing of the ensemble: one process per GPU (torch.distributed, NCCL over
NVLink/NVSwitch), particles sharded by contiguous ranges, model data replicated.

Particles are independent (src/integrator.py:105-120 reads only column i; the accept test
src/HMC.py:173-176 is per particle), so the data path needs NO collective.  Only the
per-iteration ensemble statistics vector of ehmc_hmc_iter (2D+3 float64 sums: n_accept,
sum acceptance probability, sum H, sum q_d, sum q_d^2) crosses GPUs, through one all-reduce
that is issued on a side stream and consumed one iteration late, so its ~10-20 us latency (and the
device-to-host copy of the reduced vector) is off the critical path.
"""
from __future__ import annotations

import math


def shard_range(numParticles, rank, worldSize):
    """Contiguous particle range [lo, hi) of `rank` (global Philox ids: particleOffset = lo)."""
    lo = rank * numParticles // worldSize
    hi = (rank + 1) * numParticles // worldSize
    return lo, hi


class StatsReducer:
    """Sums the statistics vectors of all ranks with one all-reduce per call.

    reduce_async(t) starts the collective on `t` in place (no-op without a process group);
    wait() blocks until the last started one is done.  Works with NCCL (CUDA tensors) and
    gloo (CPU tensors; used by the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.work = None

    def reduce_async(self, stats):
        if self.enabled:
            self.work = self.dist.all_reduce(stats, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
        return stats

    def reduce(self, stats):
        """In-place all-reduce ordered on the CURRENT stream (the host does not block with NCCL)."""
        if self.enabled:
            self.dist.all_reduce(stats, op=self.dist.ReduceOp.SUM, group=self.group)
        return stats

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None


def exchange_handles(local, group=None, device=None):
    """All-gather one fixed-size opaque handle (bytes) per rank; returns the list ordered by rank.  Works with
    NCCL (pass the CUDA `device`) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    mine = torch.frombuffer(bytearray(local), dtype=torch.uint8).clone()
    if device is not None:
        mine = mine.to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


_comms = {}


def ensemble_comm(ctx, group=None, device=None):
    """The ehmc communicator (peer-memory mailboxes) of this process for `group`, created and connected on first
    use: the in-kernel all-reduce of the fused ensemble run (ehmc_hmc_run_ensemble).  None for a single rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    key = (id(ctx), id(group))
    comm = _comms.get(key)
    if comm is None:
        from . import _lib

        nccl = dist.get_backend(group) == "nccl"
        comm = _lib.Comm(ctx, dist.get_rank(group), dist.get_world_size(group))
        comm.connect(exchange_handles(comm.local_handle(), group, device if nccl else None))
        dist.barrier(group)  # every rank has mapped every mailbox before anyone launches
        _comms[key] = comm
    return comm


class StepSizeAdapter:
    """Ensemble-based step-size adaptation (build-defined; the reference has none, SURVEY.md
    section 2).  The acceptance statistic is an average over the WHOLE ensemble (all GPUs), so a
    plain Robbins-Monro update of log(h) towards the target acceptance is already low-noise:
        log h <- log h + clip(gain_k * (mean acceptance probability - target)),  gain_k = gain0 / k^kappa
    Every rank applies the same update to the same all-reduced numbers, so step sizes stay
    identical on all ranks without a broadcast."""

    def __init__(self, stepSize, target=0.8, gain0=1.5, kappa=0.5, minStep=1e-6, maxStep=1e3, maxMove=0.7):
        self.logh = math.log(stepSize)
        self.target = float(target)
        self.gain0 = float(gain0)
        self.kappa = float(kappa)
        self.k = 0
        self.lo, self.hi = math.log(minStep), math.log(maxStep)
        self.maxMove = float(maxMove)  # cap of one update in log space

    @property
    def stepSize(self):
        return math.exp(self.logh)

    def update(self, meanAcceptProb):
        if not math.isfinite(meanAcceptProb):
            meanAcceptProb = 0.0
        self.k += 1
        move = self.gain0 / self.k**self.kappa * (meanAcceptProb - self.target)
        self.logh += min(max(move, -self.maxMove), self.maxMove)
        self.logh = min(max(self.logh, self.lo), self.hi)
        return self.stepSize


class MassAdapter:
    """Ensemble-based diagonal mass adaptation (build-defined; BASELINE.json north_star: "step-size/mass adaptation
    moments").  Consumes the all-reduced moments sum q_d, sum q_d^2 of a window of iterations -- the same numbers on
    every rank, so every rank derives the same masses without a broadcast -- and returns per-dimension scales
    s_d = sqrt(var_d): the coordinates are rescaled to q / s (unit marginal variances), which is HMC with the diagonal
    mass matrix M_d = mass / s_d^2 (Stan's "diag_e" metric, M^-1 = var).  Shrunk towards 1 like Stan's estimator,
    var <- n / (n + 5) var + 1e-3 * 5 / (n + 5), n = particles * iterations of the window."""

    def __init__(self, numDimensions, minScale=1e-8, maxScale=1e8):
        self.numDimensions = int(numDimensions)
        self.minScale, self.maxScale = float(minScale), float(maxScale)

    def scales(self, mean, var, count):
        """mean, var: per-dimension moments over `count` samples (particles x iterations, all ranks)."""
        import numpy as np

        var = np.asarray(var, dtype=np.float64)
        n = float(count)
        v = n / (n + 5.0) * var + 1e-3 * 5.0 / (n + 5.0)
        s = np.sqrt(np.where(np.isfinite(v) & (v > 0), v, 1.0))
        return np.clip(s, self.minScale, self.maxScale)


def mass_windows(adaptIterations, numWindows=3, first=0.15, last=0.1):
    """Schedule of a mass-adapting warm-up of adaptIterations iterations, Stan-like: an initial step-size-only stretch
    (`first` of the phase), numWindows windows of doubling length whose moments each give a new mass at their end,
    and a final stretch (`last`) that re-tunes the step size for the final mass.  Returns a list of
    (iterations, updateMassAfter) with the iterations summing to adaptIterations; phases too short to split keep the
    step-size adaptation only."""
    n = int(adaptIterations)
    if n <= 0:
        return []
    if n < 20 or numWindows < 1:
        return [(n, False)]
    n0 = max(1, int(round(first * n)))
    n1 = max(1, int(round(last * n)))
    rest = n - n0 - n1
    unit = rest / float(2**numWindows - 1)
    w = [max(1, int(round(unit * 2**k))) for k in range(numWindows)]
    w[-1] += rest - sum(w)
    if w[-1] < 1:
        return [(n, False)]
    return [(n0, False)] + [(k, True) for k in w] + [(n1, False)]


def unpack_stats(stats, numDimensions, numParticles):
    """Named view of the 2D+3 statistics vector of ehmc_hmc_iter (summed over ranks) for an
    ensemble of numParticles particles in total."""
    D, P = numDimensions, float(numParticles)
    s = [float(x) for x in stats[:3]]
    mean = stats[3:3 + D] / P
    var = stats[3 + D:3 + 2 * D] / P - mean * mean
    return dict(acceptRate=s[0] / P, meanAcceptProb=s[1] / P, meanH=s[2] / P, mean=mean, var=var)
