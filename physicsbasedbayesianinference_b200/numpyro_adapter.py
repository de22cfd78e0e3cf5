"""NumPyro-model -> potential-family adapter.

The reference intends potentials built from a NumPyro model through
``-numpyro.infer.util.log_density(model, args, kwargs, params)[0]`` and ``jax.grad`` of it
(samples/NumpyroExamples/CoinToss/CoinTossExample.py:75-107).  A fused CUDA trajectory kernel
cannot trace an arbitrary model, so the adapter RECOGNISES the model families the engine has
kernels for and raises for everything else.  NumPyro / JAX are optional imports (absent in the
build image); recognition works on a declarative spec so it is testable without them.
"""
from __future__ import annotations

import numpy as np

from .potential import CoinTossPotential, FunnelPotential, GaussianPotential, HarmonicPotential, LogisticPotential


def potentialFromSpec(spec):
    """spec: dict(family=..., **parameters) -> potential descriptor.

    families: "normal_iid" (scale per dim), "mvn" (mean, cov | precision),
    "funnel" (numDimensions, sigmaV), "coin_toss" (observations | successes, trials),
    "logistic_regression" (X, y, priorScale)."""
    fam = spec.get("family")
    if fam == "normal_iid":
        scale = np.atleast_1d(np.asarray(spec["scale"], dtype=np.float64))
        return HarmonicPotential(1.0 / scale**2)
    if fam == "mvn":
        if "precision" in spec:
            return GaussianPotential(precision=spec["precision"], mean=spec.get("mean"))
        return GaussianPotential(cov=spec["cov"], mean=spec.get("mean"))
    if fam == "funnel":
        return FunnelPotential(int(spec["numDimensions"]), float(spec.get("sigmaV", 3.0)))
    if fam == "coin_toss":  # Uniform(0, 1) priors, Bernoulli observations (CoinToss.py:21-25)
        if "observations" in spec:
            return CoinTossPotential.fromObservations(*spec["observations"])
        return CoinTossPotential(spec["successes"], spec["trials"])
    if fam == "logistic_regression":
        return LogisticPotential(spec["X"], spec["y"], float(spec.get("priorScale", 1.0)),
                                 precision=spec.get("precision", "fp32"))
    raise NotImplementedError(
        f"model family {fam!r} has no fused CUDA kernel; supported: normal_iid, mvn, funnel, coin_toss, logistic_regression")


def potentialFromNumpyroModel(model, model_args=(), model_kwargs=None):
    """Traces a NumPyro model once and maps it to a family (needs numpyro + jax).  Recognised:
    a single MultivariateNormal / Normal latent site without observations, and a Bernoulli-logit
    likelihood ``obs ~ Bernoulli(logits = X @ theta)`` with a Normal(0, s) prior on theta."""
    try:
        import numpyro  # noqa: F401
        from numpyro import handlers
        import numpyro.distributions as dist
        import jax
    except ImportError as e:  # pragma: no cover - numpyro is not in the build image
        raise ImportError("potentialFromNumpyroModel needs numpyro and jax; use potentialFromSpec instead") from e
    tr = handlers.trace(handlers.seed(model, jax.random.PRNGKey(0))).get_trace(*model_args, **(model_kwargs or {}))
    latent = [s for s in tr.values() if s["type"] == "sample" and not s["is_observed"]]
    observed = [s for s in tr.values() if s["type"] == "sample" and s["is_observed"]]
    if latent and len(latent) == len(observed) and all(
            isinstance(s["fn"], dist.Uniform) and float(s["fn"].low) == 0.0 and float(s["fn"].high) == 1.0 for s in latent) and all(
            isinstance(getattr(s["fn"], "base_dist", s["fn"]), (dist.BernoulliProbs,)) for s in observed):
        # the coin-toss sample: site k observes Bernoulli(latent k)
        return potentialFromSpec(dict(family="coin_toss", observations=[np.asarray(s["value"]) for s in observed]))
    if len(latent) == 1 and not observed:
        d = latent[0]["fn"]
        if isinstance(d, dist.MultivariateNormal):
            return potentialFromSpec(dict(family="mvn", mean=np.asarray(d.mean), cov=np.asarray(d.covariance_matrix)))
        base = getattr(d, "base_dist", d)
        if isinstance(base, dist.Normal) and np.allclose(np.asarray(base.loc), 0.0):
            return potentialFromSpec(dict(family="normal_iid", scale=np.broadcast_to(np.asarray(base.scale), latent[0]["value"].shape)))
    raise NotImplementedError("this NumPyro model does not match a family with a fused CUDA kernel; "
                              "describe it with potentialFromSpec(dict(family=...)) if it is one of them")
