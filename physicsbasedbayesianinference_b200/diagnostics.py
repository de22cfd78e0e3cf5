"""Effective sample size of ensemble chains (build-defined: the reference has no ESS).

Estimator: for one coordinate, the chains are centred on the GRAND mean over chains and draws (they
are exchangeable: same target, independent particles -- centring every chain on its own mean would
hide exactly the slow component a poorly mixing chain has not traversed yet), the autocovariance is
computed by FFT per chain and AVERAGED over the traced chains, then
Geyer's initial-positive-sequence truncation gives the integrated autocorrelation time tau;
ESS = (number of draws) x (number of chains) / tau.  Runs on the tensors' device (torch.fft)."""
from __future__ import annotations

import math


def ess(trace):
    """trace: torch tensor (S, C) -- S draws of C chains of ONE coordinate.  Returns total ESS."""
    import torch

    x = trace.to(torch.float64)
    S, C = x.shape
    xc = x - x.mean()
    n = 1 << (2 * S - 1).bit_length()
    f = torch.fft.rfft(xc, n=n, dim=0)
    acov = torch.fft.irfft(f * f.conj(), n=n, dim=0)[:S] / S
    acov = acov.mean(dim=1).cpu()
    if float(acov[0]) <= 0:
        return float(S * C)
    rho = (acov / acov[0]).tolist()
    tau = -1.0
    t = 0
    while t + 1 < S:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        tau += 2.0 * pair
        t += 2
    tau = max(tau, 1.0 / math.log10(max(S, 10)))
    return float(S * C / tau)


def ess_min_over_dims(trace, numParticlesTotal=None):
    """trace: (D, C, S) positions of C traced particles over S iterations.  Returns
    (min over dimensions of the ESS of the traced chains, the same scaled to the whole ensemble)."""
    D, C, S = trace.shape
    vals = [ess(trace[d].transpose(0, 1)) for d in range(D)]
    m = min(vals)
    scale = (numParticlesTotal / C) if numParticlesTotal else 1.0
    return m, m * scale
