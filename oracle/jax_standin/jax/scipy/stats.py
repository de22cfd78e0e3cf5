"""jax.scipy.stats stand-in (only multivariate_normal.pdf/logpdf are used by
the reference's tests: src/tests/test_HMC.py:48-49,123-124)."""
from scipy.stats import multivariate_normal  # noqa: F401
