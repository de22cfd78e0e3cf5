"""jax.numpy stand-in: re-export NumPy (float64).  Test infrastructure only."""
from numpy import *  # noqa: F401,F403
from numpy import linalg  # noqa: F401
import numpy as _np

pi = _np.pi
