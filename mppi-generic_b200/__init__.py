"""mppi-generic_b200 — Blackwell-native MPPI rollout-and-reduce engine behind the reference's plugin surface.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C-ABI, built into ``libmppi_b200.so``) and
``host.py`` (ctypes mirror of the reference's host classes). The directory name carries a hyphen, so import it through
the top-level shim module ``mppi_generic_b200`` (same package object).
"""
from .host import *  # noqa: F401,F403
from . import host, workloads  # noqa: F401
