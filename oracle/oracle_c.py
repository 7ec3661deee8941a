"""ctypes wrapper of oracle/scdc_oracle.c (CPU ORACLE — test infrastructure, not the product path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "scdc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/liboracle.so"] + (["-B"] if force else []))
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def perm_counts(problem, perm_ct, perm_cond=None, score_type=0, return_distributions=False, nthreads=0, de_scale=None):
    """problem: any object with the scdc_problem fields (e.g. scdiffcom_b200.PermutationProblem).
    perm_ct / perm_cond: uint8 [iterations, N]. Returns (counts float64 [R, ncols] with NaN where observed is NaN,
    distributions [iterations, ncols_d, R] or None). de_scale: optional float64 [iterations, R] output (differential mode):
    pass-1 score of cond2 + of cond1, the scale of the DE column's rounding error."""
    lib = load()
    N, G, T, Cn = problem.n_cells, problem.n_genes, problem.n_celltypes, problem.n_conditions
    nL, nR = problem.max_nL, problem.max_nR
    ncols = 1 if Cn == 1 else 3 + nL + nR
    R = int(np.asarray(problem.row_lri).shape[0])
    perm_ct = np.ascontiguousarray(perm_ct, dtype=np.uint8)
    iters = perm_ct.shape[0]
    assert perm_ct.shape == (iters, N)
    if Cn == 2:
        perm_cond = np.ascontiguousarray(perm_cond, dtype=np.uint8)
        assert perm_cond.shape == (iters, N)
    sub_gene = _i32(np.asarray(problem.sub_gene).reshape(-1, nL + nR))
    obs = np.asfortranarray(np.asarray(problem.observed, dtype=np.float64).reshape(R, ncols))
    counts = np.zeros((R, ncols), dtype=np.uint64, order="F")
    dist = np.zeros((iters, 3 if Cn == 2 else 1, R), dtype=np.float64) if return_distributions else None
    colptr, rowidx = _i32(problem.colptr), _i32(problem.rowidx)
    values = np.ascontiguousarray(problem.values, dtype=np.float64)
    ct, rl, re_, rr = _i32(problem.ct_code), _i32(problem.row_lri), _i32(problem.row_emit), _i32(problem.row_recv)
    lib.oracle_perm_counts.argtypes = [C.c_int] * 8 + [C.c_void_p] * 9 + [C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                                                          C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    rc = lib.oracle_perm_counts(N, G, T, Cn, sub_gene.shape[0], R, nL, nR, _p(colptr), _p(rowidx), _p(values), _p(ct),
                                _p(sub_gene), _p(rl), _p(re_), _p(rr), _p(obs), iters,
                                int(score_type), _p(perm_cond), _p(perm_ct), _p(counts), _p(dist), int(nthreads),
                                _p(de_scale))
    if rc:
        raise MemoryError("oracle_perm_counts: allocation failed")
    out = counts.astype(np.float64)
    out[np.isnan(obs)] = np.nan
    return out, dist


def group_means(problem, group, detection=False):
    lib = load()
    K = problem.n_conditions * problem.n_celltypes
    out = np.zeros((K, problem.n_genes), dtype=np.float64, order="F")
    colptr, rowidx = _i32(problem.colptr), _i32(problem.rowidx)
    values = np.ascontiguousarray(problem.values, dtype=np.float64)
    group = np.ascontiguousarray(group, dtype=np.uint8)
    lib.oracle_group_means.argtypes = [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
    rc = lib.oracle_group_means(problem.n_cells, problem.n_genes, K, _p(colptr), _p(rowidx), _p(values), _p(group),
                                1 if detection else 0, _p(out))
    if rc:
        raise MemoryError
    return out


def philox_labels(problem, n, seed=42, perm_offset=0):
    """CPU restatement of the device label generator -> (cond uint8 [n, N] or None, ct uint8 [n, N])."""
    lib = load()
    N, T, Cn = problem.n_cells, problem.n_celltypes, problem.n_conditions
    ct_out = np.zeros((n, N), dtype=np.uint8)
    cond_out = np.zeros((n, N), dtype=np.uint8) if Cn == 2 else None
    ct, cd = _i32(problem.ct_code), _i32(problem.cond_code)
    lib.oracle_philox_labels.argtypes = [C.c_int] * 3 + [C.c_void_p] * 2 + [C.c_uint64, C.c_int64, C.c_int64, C.c_void_p,
                                                                         C.c_void_p]
    rc = lib.oracle_philox_labels(N, T, Cn, _p(ct), _p(cd), int(seed) & (2**64 - 1), int(perm_offset), int(n),
                                  _p(cond_out), _p(ct_out))
    if rc:
        raise MemoryError
    return cond_out, ct_out
