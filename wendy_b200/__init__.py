"""wendy_b200: B200-native (sm_100a) approximate-integration hot path of jobovy/wendy."""
from .wendy import nbody, energy, momentum, potential, argsort, trim, ApproxState  # noqa: F401
from . import ic, multi  # noqa: F401
