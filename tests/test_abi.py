"""CPU tests of the C-ABI shared library: it loads, exports exactly what include/scdc.h declares, validates arguments
up front, and refuses to compute without a B200 (no CPU fallback). No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from scdiffcom_b200 import _lib, engine
from tests import helpers  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "scdc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scdc_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert sorted(syms) == sorted(_lib.EXPORTED)
    for s in syms:
        assert hasattr(lib, s), s
    ver = re.search(r'SCDC_VERSION_STRING "([^"]+)"', open(os.path.join(ROOT, "include", "scdc.h")).read()).group(1)
    assert _lib.load().scdc_version().decode() == ver


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors in _lib.py against the C compiler's view of include/scdc.h: size and offset of every field."""
    import subprocess
    structs = {"scdc_problem": _lib.Problem, "scdc_options": _lib.Options, "scdc_stats": _lib.Stats,
               "scdc_result": _lib.Result, "scdc_observed": _lib.Observed}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "scdc.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = {}
    for ln in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, fname, val = ln.split()
        got[(cname, fname)] = int(val)
    for cname, cls in structs.items():
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)


def _tiny_problem():
    # 6 cells, 2 genes, 2 cell types, detection mode, 1 LRI, 1 row
    return engine.PermutationProblem(
        n_cells=6, n_genes=2, n_celltypes=2, n_conditions=1, max_nL=1, max_nR=1,
        colptr=np.array([0, 3, 5]), rowidx=np.array([0, 2, 4, 1, 5]), values=np.array([1.0, 2.0, 3.0, 4.0, 5.0]),
        ct_code=np.array([0, 1, 0, 1, 0, 1]), cond_code=None, sub_gene=np.array([[0, 1]]),
        row_lri=np.array([0]), row_emit=np.array([0]), row_recv=np.array([1]), observed=np.array([[1.0]]))


def test_argument_validation_happens_before_any_device_work():
    bad = _tiny_problem()
    bad.rowidx = np.array([0, 2, 9, 1, 5])
    with pytest.raises(_lib.ScdcError, match="SCDC_EINVAL.*rowidx out of range"):
        engine.perm_counts(bad, 10)
    bad = _tiny_problem()
    bad.rowidx = np.array([2, 0, 4, 1, 5])
    with pytest.raises(_lib.ScdcError, match="ascending"):
        engine.perm_counts(bad, 10)
    bad = _tiny_problem()
    bad.ct_code = np.array([0, 0, 0, 0, 0, 0])
    with pytest.raises(_lib.ScdcError, match="has no cells"):
        engine.perm_counts(bad, 10)
    bad = _tiny_problem()
    bad.sub_gene = np.array([[0, 7]])
    with pytest.raises(_lib.ScdcError, match="sub_gene"):
        engine.perm_counts(bad, 10)
    with pytest.raises(_lib.ScdcError, match="iterations"):
        engine.perm_counts(_tiny_problem(), 0)
    with pytest.raises(_lib.ScdcError, match="perm_cond given in detection mode"):
        engine.perm_counts(_tiny_problem(), 2, perm_cond=np.zeros((2, 6), np.uint8), perm_ct=np.zeros((2, 6), np.uint8))


def test_no_gpu_means_error_not_fallback():
    if engine.device_count() > 0:
        pytest.skip("a B200 is visible")
    with pytest.raises(_lib.ScdcError, match="SCDC_ENOGPU"):
        engine.perm_counts(_tiny_problem(), 10)
    with pytest.raises(_lib.ScdcError, match="SCDC_ENOGPU"):
        engine.Engine(_tiny_problem())


def test_rcall_wrapper_compiles_against_stub_r_headers(tmp_path):
    """src/scdc_rcall.c is the only file that includes Rinternals.h; R is not installed, so compile-check it against the
    minimal stub headers in tests/r_stub/."""
    import subprocess
    src = os.path.join(ROOT, "src", "scdc_rcall.c")
    if not os.path.exists(src):
        pytest.skip("src/scdc_rcall.c not present yet")
    out = tmp_path / "scdc_rcall.o"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-c", src, "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "r_stub"), "-o", str(out)])
    assert out.exists()
