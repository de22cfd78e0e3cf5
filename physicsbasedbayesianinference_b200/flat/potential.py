"""Flat-module shim: ``from potential import ...`` as in the reference's src/potential.py."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from physicsbasedbayesianinference_b200.potential import *  # noqa: E402,F401,F403
from physicsbasedbayesianinference_b200 import potential as _m  # noqa: E402

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
