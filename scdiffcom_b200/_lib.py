"""ctypes binding of the C-ABI in include/scdc.h (libscdc.so, built in-tree by `__graft_entry__.build()`).

There is no Python/CPU implementation behind this module: if the shared library is missing, or no B200 is visible,
every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscdc.so")

SCDC_OK, SCDC_EINVAL, SCDC_ENOGPU, SCDC_ECUDA, SCDC_ENCCL, SCDC_ENOMEM, SCDC_EINTERRUPT = range(7)
SCDC_COUNT_NA = 2**64 - 1
PRECISIONS = {"fp64": 0, "fp64_exact": 0, "fp32_fast": 1}
HAS_FP32_FAST = True   # SCDC_FP32_FAST is implemented by this library version (bench.py reports it under an extra key)
_ERRNAMES = {1: "SCDC_EINVAL", 2: "SCDC_ENOGPU", 3: "SCDC_ECUDA", 4: "SCDC_ENCCL", 5: "SCDC_ENOMEM", 6: "SCDC_EINTERRUPT"}

# every symbol include/scdc.h declares (tests/test_abi.py checks header <-> library <-> this list)
EXPORTED = (
    "scdc_perm_counts", "scdc_create", "scdc_run", "scdc_release", "scdc_group_means", "scdc_draw_labels",
    "scdc_set_rows", "scdc_observed_pass", "scdc_trim", "scdc_cci_template", "scdc_merge_pvalues", "scdc_bh_adjust",
    "scdc_selftest_de_filter", "scdc_last_error", "scdc_device_count", "scdc_version",
)


class ScdcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {message}")
        self.code = code


class Problem(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int32), ("n_genes", C.c_int32), ("n_celltypes", C.c_int32), ("n_conditions", C.c_int32),
        ("n_lri", C.c_int32), ("n_rows", C.c_int32), ("max_nL", C.c_int32), ("max_nR", C.c_int32),
        ("colptr", C.c_void_p), ("rowidx", C.c_void_p), ("values", C.c_void_p), ("ct_code", C.c_void_p),
        ("cond_code", C.c_void_p), ("sub_gene", C.c_void_p), ("row_lri", C.c_void_p), ("row_emit", C.c_void_p),
        ("row_recv", C.c_void_p), ("observed", C.c_void_p),
    ]


class Options(C.Structure):
    _fields_ = [
        ("iterations", C.c_int64), ("seed", C.c_uint64), ("perm_offset", C.c_int64), ("score_type", C.c_int32),
        ("n_gpus", C.c_int32), ("precision", C.c_int32), ("batch_size", C.c_int32), ("perm_cond", C.c_void_p),
        ("perm_ct", C.c_void_p), ("return_distributions", C.c_int32), ("device_base", C.c_int32), ("stream", C.c_void_p),
        ("interrupt", C.c_void_p), ("interrupt_user", C.c_void_p), ("rel_tol", C.c_double),
    ]


INTERRUPT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)   # int (*interrupt)(void* user)


class Stats(C.Structure):
    _fields_ = [
        ("ms_total", C.c_double), ("ms_upload", C.c_double), ("ms_labels", C.c_double), ("ms_reduce", C.c_double),
        ("ms_score", C.c_double), ("ms_merge", C.c_double), ("ms_device", C.c_double), ("reduce_bytes", C.c_double),
        ("reduce_wavefronts", C.c_double),
        ("reduce_launches", C.c_int64), ("kernel_launches", C.c_int64), ("batch_size", C.c_int64),
        ("n_gpus", C.c_int32), ("reserved", C.c_int32),
        ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double), ("p2p_bytes", C.c_double),
    ]

    def as_dict(self):
        d = {name: getattr(self, name) for name, _ in self._fields_ if name != "reserved"}
        d["used_nccl"] = int(self.reserved)   # multi-GPU calls: 1 = NCCL all-reduce of the counters, 0 = peer copies + adds
        return d


class Result(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("distributions", C.c_void_p), ("counts_device", C.c_void_p),
                ("band_counts", C.c_void_p), ("band_device", C.c_void_p), ("stats", Stats)]


class Observed(C.Structure):
    _fields_ = [("expression", C.c_void_p), ("detection_rate", C.c_void_p), ("score", C.c_void_p),
                ("is_expressed", C.c_void_p)]


_lib = None


def load():
    """Load libscdc.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). scdiffcom_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.scdc_last_error.restype = C.c_char_p
    lib.scdc_version.restype = C.c_char_p
    lib.scdc_device_count.restype = C.c_int
    lib.scdc_perm_counts.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(Result)]
    lib.scdc_create.argtypes = [C.POINTER(Problem), C.c_int32, C.POINTER(C.c_void_p)]
    lib.scdc_run.argtypes = [C.c_void_p, C.POINTER(Options), C.POINTER(Result)]
    lib.scdc_release.argtypes = [C.c_void_p]
    lib.scdc_release.restype = None
    lib.scdc_group_means.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scdc_draw_labels.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.scdc_set_rows.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scdc_observed_pass.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                       C.c_int32, C.POINTER(Observed)]
    lib.scdc_trim.argtypes = [C.c_int32]
    lib.scdc_cci_template.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 7
    lib.scdc_bh_adjust.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scdc_merge_pvalues.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scdc_selftest_de_filter.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    _lib = lib
    return lib


def check(rc):
    if rc != SCDC_OK:
        raise ScdcError(rc, load().scdc_last_error().decode("utf-8", "replace"))
