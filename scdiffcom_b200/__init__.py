"""scdiffcom_b200 — B200-native (sm_100a) engine for scDiffCom's permutation test.

Only the hot path of `run_interaction_analysis()` is here: `run_stat_analysis()` (reference
R/utils_permutation.R:1-499) backed by hand-written CUDA kernels behind the C-ABI of `include/scdc.h`.
Importing the package does not need a GPU; computing anything does (there is no CPU fallback).
"""
from ._lib import ScdcError, LIB_PATH  # noqa: F401
from .engine import Engine, PermutationProblem, device_count, perm_counts, trim, version  # noqa: F401
from .permutation import run_stat_analysis  # noqa: F401

__all__ = ["Engine", "PermutationProblem", "ScdcError", "device_count", "perm_counts", "run_stat_analysis", "trim", "version"]
