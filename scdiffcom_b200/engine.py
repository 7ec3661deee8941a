"""Thin object layer over the C-ABI: marshals integer-coded NumPy arrays into `scdc_problem` and runs the B200 engine.

The arrays are exactly what SURVEY.md section 8b says the R glue passes through `.Call` (see include/scdc.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import Options, Problem, Result, ScdcError, check  # noqa: F401

SCORE_TYPES = {"geometric_mean": 0, "arithmetic_mean": 1}


def _arr(x, dtype):
    if x is None:
        return None
    return np.ascontiguousarray(x, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class PermutationProblem:
    """analysis_inputs + sub_cci_template_iter + observed vectors, integer-encoded (include/scdc.h: scdc_problem)."""
    n_cells: int
    n_genes: int
    n_celltypes: int
    n_conditions: int
    max_nL: int
    max_nR: int
    colptr: np.ndarray      # int32 [G+1]
    rowidx: np.ndarray      # int32 [nnz]
    values: np.ndarray      # float64 [nnz]
    ct_code: np.ndarray     # int32 [N]
    cond_code: np.ndarray | None  # int32 [N] or None (detection mode)
    sub_gene: np.ndarray    # int32 [L, nL+nR], -1 = NA
    row_lri: np.ndarray     # int32 [R_sub]
    row_emit: np.ndarray
    row_recv: np.ndarray
    observed: np.ndarray    # float64 [R_sub, ncols] (any layout; passed column-major)
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_rows(self):
        return int(self.row_lri.shape[0])

    @property
    def n_lri(self):
        return int(self.sub_gene.shape[0])

    @property
    def ncols(self):
        return 1 if self.n_conditions == 1 else 3 + self.max_nL + self.max_nR

    @property
    def nnz(self):
        return int(self.colptr[-1])

    def c_struct(self) -> Problem:
        self.colptr = _arr(self.colptr, np.int32)
        self.rowidx = _arr(self.rowidx, np.int32)
        self.values = _arr(self.values, np.float64)
        self.ct_code = _arr(self.ct_code, np.int32)
        self.cond_code = _arr(self.cond_code, np.int32)
        self.sub_gene = _arr(np.asarray(self.sub_gene).reshape(-1, self.max_nL + self.max_nR), np.int32)
        self.row_lri = _arr(self.row_lri, np.int32)
        self.row_emit = _arr(self.row_emit, np.int32)
        self.row_recv = _arr(self.row_recv, np.int32)
        obs = np.asarray(self.observed, dtype=np.float64).reshape(self.n_rows, self.ncols)
        obs_f = np.asfortranarray(obs)
        self.observed = obs_f     # column-major once: later calls on the same problem do not copy it again
        self._keep = [obs_f]
        p = Problem()
        p.n_cells, p.n_genes, p.n_celltypes, p.n_conditions = self.n_cells, self.n_genes, self.n_celltypes, self.n_conditions
        p.n_lri, p.n_rows, p.max_nL, p.max_nR = self.n_lri, self.n_rows, self.max_nL, self.max_nR
        p.colptr, p.rowidx, p.values = _ptr(self.colptr), _ptr(self.rowidx), _ptr(self.values)
        p.ct_code, p.cond_code, p.sub_gene = _ptr(self.ct_code), _ptr(self.cond_code), _ptr(self.sub_gene)
        p.row_lri, p.row_emit, p.row_recv = _ptr(self.row_lri), _ptr(self.row_emit), _ptr(self.row_recv)
        p.observed = _ptr(obs_f)
        return p


def _options(iterations, seed, score_type, perm_offset, n_gpus, batch_size, perm_cond, perm_ct, return_distributions,
             n_cells, keep, device_base=0, stream=None, interrupt=None, precision="fp64", rel_tol=0.0):
    o = Options()
    o.iterations = int(iterations)
    o.seed = int(seed) & (2**64 - 1)
    o.perm_offset = int(perm_offset)
    o.score_type = SCORE_TYPES[score_type] if isinstance(score_type, str) else int(score_type)
    o.n_gpus = int(n_gpus)
    o.precision = _lib.PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    o.rel_tol = float(rel_tol)
    o.batch_size = int(batch_size)
    for name, lab in (("perm_cond", perm_cond), ("perm_ct", perm_ct)):
        if lab is not None:
            a = np.ascontiguousarray(lab, dtype=np.uint8)
            if a.shape != (int(iterations), n_cells):
                raise ValueError(f"{name} must have shape (iterations, n_cells) = ({iterations}, {n_cells}), got {a.shape}")
            keep.append(a)
            setattr(o, name, _ptr(a))
    o.return_distributions = 1 if return_distributions else 0
    o.device_base = int(device_base)
    o.stream = C.c_void_p(int(stream)) if stream else None
    if interrupt is not None:   # callable() -> truthy to stop; polled by the library on this thread between batches
        cb = _lib.INTERRUPT_FN(lambda _user: 1 if interrupt() else 0)
        keep.append(cb)
        o.interrupt = C.cast(cb, C.c_void_p)
    return o


def _finish(problem, counts_f, dist, stats, band_f=None, as_float=True):
    if not as_float:   # raw uint64 counters (SCDC_COUNT_NA = NA): what the .Call wrapper receives, no float copies
        return {"counts": None, "counts_raw": counts_f, "band_counts": None, "band_raw": band_f, "distributions": dist,
                "stats": stats.as_dict()}
    na = counts_f == np.uint64(_lib.SCDC_COUNT_NA)
    counts = counts_f.astype(np.float64)
    counts[na] = np.nan
    band = None
    if band_f is not None:
        band = band_f.astype(np.float64)
        band[na] = np.nan
    return {"counts": counts, "counts_raw": counts_f, "band_counts": band, "distributions": dist, "stats": stats.as_dict()}


def device_count() -> int:
    return int(_lib.load().scdc_device_count())


def trim(device: int = -1) -> None:
    """Return the engine's cached device memory (stream-ordered pool) to the driver."""
    check(_lib.load().scdc_trim(int(device)))


def version() -> str:
    return _lib.load().scdc_version().decode()


def perm_counts(problem: PermutationProblem, iterations, seed=42, score_type="geometric_mean", n_gpus=1, perm_offset=0,
                batch_size=0, perm_cond=None, perm_ct=None, return_distributions=False, device_base=0, interrupt=None,
                precision="fp64", rel_tol=0.0, as_float=True):
    """One-shot `scdc_perm_counts` with HOST buffers (what the `.Call` wrapper does). Returns dict(counts [R_sub, ncols]
    float64 with NaN = NA, band_counts (precision="fp32_fast": comparisons inside the tolerance band) or None,
    distributions or None, stats). as_float=False skips the float64 copies and returns only `counts_raw` (uint64 with
    SCDC_COUNT_NA), i.e. exactly what the C-ABI call hands to the `.Call` wrapper."""
    lib = _lib.load()
    keep = []
    p = problem.c_struct()
    o = _options(iterations, seed, score_type, perm_offset, n_gpus, batch_size, perm_cond, perm_ct, return_distributions,
                 problem.n_cells, keep, device_base=device_base, interrupt=interrupt, precision=precision, rel_tol=rel_tol)
    counts = np.zeros((problem.n_rows, problem.ncols), dtype=np.uint64, order="F")
    res = Result()
    res.counts = _ptr(counts)
    band = None
    if o.precision != 0:
        band = np.zeros((problem.n_rows, problem.ncols), dtype=np.uint64, order="F")
        res.band_counts = _ptr(band)
    dist = None
    if return_distributions:
        ncd = 3 if problem.n_conditions == 2 else 1
        dist = np.zeros((int(iterations), ncd, problem.n_rows), dtype=np.float64)
        res.distributions = _ptr(dist)
    check(lib.scdc_perm_counts(C.byref(p), C.byref(o), C.byref(res)))
    return _finish(problem, counts, dist, res.stats, band, as_float=as_float)


def cci_template(n_celltypes, n_conditions, lri_order, ct_code, cond_code=None):
    """`scdc_cci_template` (create_cci_template + add_cell_number, R/utils_cci.R:1-169, integer-coded): returns
    (emitter, receiver, lri_idx) int32 [T*T*L] and (ncells_emitter, ncells_receiver) int32 [C, T*T*L]."""
    lib = _lib.load()
    order = _arr(lri_order, np.int32)
    T, Cn, L = int(n_celltypes), int(n_conditions), len(order)
    n = T * T * L
    em, rc, li = (np.empty(n, dtype=np.int32) for _ in range(3))
    ne, nr = np.empty((Cn, n), dtype=np.int32), np.empty((Cn, n), dtype=np.int32)
    ct, cd = _arr(ct_code, np.int32), _arr(cond_code, np.int32)
    check(lib.scdc_cci_template(T, Cn, L, _ptr(order), len(ct), _ptr(ct), _ptr(cd), _ptr(em), _ptr(rc), _ptr(li), _ptr(ne), _ptr(nr)))
    return em, rc, li, ne, nr


def merge_pvalues(n_full, tested_rows, counts_raw, iterations, sub_gene=None, lri_idx=None, max_nL=1, max_nR=1):
    """`scdc_merge_pvalues` (R/utils_permutation.R:215-259 + 427-469): counts_raw uint64 [R_sub, ncols] (SCDC_COUNT_NA = NA) ->
    p-values float64 [n_full, ncols] with 1 outside the tested rows and NaN for absent subunits."""
    lib = _lib.load()
    counts = np.asfortranarray(counts_raw, dtype=np.uint64)
    n_tested, ncols = counts.shape
    rows = _arr(tested_rows, np.int64)
    out = np.empty((int(n_full), ncols), dtype=np.float64, order="F")
    sg = None if sub_gene is None else _arr(sub_gene, np.int32)
    li = None if lri_idx is None else _arr(lri_idx, np.int32)
    check(lib.scdc_merge_pvalues(int(n_full), n_tested, _ptr(rows), ncols, _ptr(counts), int(iterations),
                                 0 if sg is None else sg.shape[0], int(max_nL), int(max_nR), _ptr(sg), _ptr(li), _ptr(out)))
    return out


def bh_adjust(p, group=None):
    """`scdc_bh_adjust`: stats::p.adjust(p, method = "BH") (NaN = NA kept), optionally inside every `group`
    (the reference's BH_P_VALUE* columns, R/utils_filtering.R:216-235)."""
    lib = _lib.load()
    p = _arr(p, np.float64)
    g = None if group is None else _arr(group, np.int32)
    out = np.empty_like(p)
    check(lib.scdc_bh_adjust(p.size, _ptr(p), _ptr(g), _ptr(out)))
    return out


class Engine:
    """Handle API: the problem stays resident in HBM on one device across runs (`scdc_create` / `scdc_run`)."""

    def __init__(self, problem: PermutationProblem, device: int = -1):
        self._lib = _lib.load()
        self.problem = problem
        self._h = C.c_void_p()
        p = problem.c_struct()
        check(self._lib.scdc_create(C.byref(p), int(device), C.byref(self._h)))

    def close(self):
        if self._h:
            self._lib.scdc_release(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, iterations, seed=42, score_type="geometric_mean", perm_offset=0, batch_size=0, perm_cond=None,
            perm_ct=None, return_distributions=False, counts_device_ptr=None, want_host_counts=True, stream=None,
            interrupt=None, precision="fp64", rel_tol=0.0, band_device_ptr=None):
        keep = []
        pr = self.problem
        o = _options(iterations, seed, score_type, perm_offset, 1, batch_size, perm_cond, perm_ct, return_distributions,
                     pr.n_cells, keep, stream=stream, interrupt=interrupt, precision=precision, rel_tol=rel_tol)
        res = Result()
        counts = band = None
        if want_host_counts:
            counts = np.zeros((pr.n_rows, pr.ncols), dtype=np.uint64, order="F")
            res.counts = _ptr(counts)
            if o.precision != 0:
                band = np.zeros((pr.n_rows, pr.ncols), dtype=np.uint64, order="F")
                res.band_counts = _ptr(band)
        if band_device_ptr is not None:
            res.band_device = C.c_void_p(int(band_device_ptr))
        if counts_device_ptr is not None:
            res.counts_device = C.c_void_p(int(counts_device_ptr))
        dist = None
        if return_distributions:
            ncd = 3 if pr.n_conditions == 2 else 1
            dist = np.zeros((int(iterations), ncd, pr.n_rows), dtype=np.float64)
            res.distributions = _ptr(dist)
        check(self._lib.scdc_run(self._h, C.byref(o), C.byref(res)))
        if counts is None:
            return {"counts": None, "distributions": dist, "stats": res.stats.as_dict()}
        return _finish(pr, counts, dist, res.stats, band)

    def group_means(self, detection_rate=False):
        """aggregate_cells on the original labels (R/utils_cci.R:441-462): [K, G] means (and detection rates)."""
        pr = self.problem
        K = pr.n_conditions * pr.n_celltypes
        means = np.zeros((K, pr.n_genes), dtype=np.float64, order="F")
        dr = np.zeros((K, pr.n_genes), dtype=np.float64, order="F") if detection_rate else None
        check(self._lib.scdc_group_means(self._h, _ptr(means), _ptr(dr)))
        return (means, dr) if detection_rate else means

    def set_rows(self, row_lri, row_emit, row_recv, observed):
        """(Re)install the tested rows + observed vectors on the resident handle (`scdc_set_rows`)."""
        pr = self.problem
        pr.row_lri, pr.row_emit, pr.row_recv = _arr(row_lri, np.int32), _arr(row_emit, np.int32), _arr(row_recv, np.int32)
        obs = np.asfortranarray(np.asarray(observed, dtype=np.float64).reshape(pr.n_rows, pr.ncols))
        pr.observed = obs
        check(self._lib.scdc_set_rows(self._h, pr.n_rows, _ptr(pr.row_lri), _ptr(pr.row_emit), _ptr(pr.row_recv), _ptr(obs)))

    def observed_pass(self, row_lri, row_emit, row_recv, score_type="geometric_mean", threshold_min_cells=5,
                      threshold_pct=0.1):
        """Observed pass over template rows on the GPU (`scdc_observed_pass`): dict(expression [n, C, nS],
        detection_rate [n, C, nS], score [n, C], is_expressed [n, C] bool)."""
        pr = self.problem
        lri, em, rc = _arr(row_lri, np.int32), _arr(row_emit, np.int32), _arr(row_recv, np.int32)
        n, Cn, nS = len(lri), pr.n_conditions, pr.max_nL + pr.max_nR
        expr = np.zeros((Cn * nS, n), dtype=np.float64)
        dr = np.zeros((Cn * nS, n), dtype=np.float64)
        score = np.zeros((Cn, n), dtype=np.float64)
        is_e = np.zeros((Cn, n), dtype=np.uint8)
        out = _lib.Observed(_ptr(expr), _ptr(dr), _ptr(score), _ptr(is_e))
        st = SCORE_TYPES[score_type] if isinstance(score_type, str) else int(score_type)
        check(self._lib.scdc_observed_pass(self._h, n, _ptr(lri), _ptr(em), _ptr(rc), float(threshold_pct),
                                           float(threshold_min_cells), st, C.byref(out)))
        return {"expression": expr.reshape(Cn, nS, n).transpose(2, 0, 1), "detection_rate": dr.reshape(Cn, nS, n).transpose(2, 0, 1),
                "score": score.T, "is_expressed": is_e.T.astype(bool)}

    def draw_labels(self, n, seed=42, perm_offset=0):
        """The label matrices the device generator produces for replicates [perm_offset, perm_offset + n)."""
        pr = self.problem
        ct = np.zeros((int(n), pr.n_cells), dtype=np.uint8)
        cond = np.zeros((int(n), pr.n_cells), dtype=np.uint8) if pr.n_conditions == 2 else None
        check(self._lib.scdc_draw_labels(self._h, int(seed) & (2**64 - 1), int(perm_offset), int(n), _ptr(cond), _ptr(ct)))
        return cond, ct
