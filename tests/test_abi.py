"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/psvo.h
declares; the ctypes struct mirrors have the C layout; argument validation does not need a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from psvo_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from psvo_b200.build import build
        build()
    return _lib.load()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "psvo.h")).read()
    return sorted(set(re.findall(r"PSVO_API\s+[\w\s\*]+?\b(psvo_\w+)\s*\(", txt)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("psvo_smc_forward", "psvo_smc_backward", "psvo_ffbsi_forward", "psvo_resample", "psvo_gather",
              "psvo_mlp_forward", "psvo_mlp_backward", "psvo_mvn_log_prob", "psvo_logsumexp", "psvo_get_sigma"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    from psvo_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (psvo_\w+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == declared_symbols()      # the ctypes table covers the header, no more no less


def test_version_and_error_text(lib):
    assert lib.psvo_version() == 200
    rc = lib.psvo_resample(None, None, 1, 4, 4, None, None)
    assert rc < 0 and b"null" in lib.psvo_last_error()


def test_struct_layout_matches_c(tmp_path):
    """compile a tiny C program against include/psvo.h and compare sizeof/offsetof with the ctypes mirrors"""
    from psvo_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "psvo.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(psvo_mlp),sizeof(psvo_dist),sizeof(psvo_params),sizeof(psvo_grads),sizeof(psvo_dims),'
                   'offsetof(psvo_dist,sigma_con),offsetof(psvo_dims,flags),offsetof(psvo_mlp,W));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.psvo_mlp), C.sizeof(_lib.psvo_dist), C.sizeof(_lib.psvo_params), C.sizeof(_lib.psvo_grads),
            C.sizeof(_lib.psvo_dims), _lib.psvo_dist.sigma_con.offset, _lib.psvo_dims.flags.offset, _lib.psvo_mlp.W.offset]
    assert got == want


def test_argument_validation_without_gpu(lib):
    from psvo_b200 import _lib
    d = _lib.psvo_dims()
    d.N, d.B, d.T, d.Dx, d.Dy = 8, 1, 2, 99, 1
    p = _lib.psvo_params()
    rc = lib.psvo_smc_forward(C.byref(d), C.byref(p), *([None] * 6), 0, *([None] * 7), 0, None)
    assert rc < 0 and b"Dx=99" in lib.psvo_last_error()


def test_product_package_does_not_import_the_oracle():
    """the oracle is test infrastructure; no module under psvo_b200/ may import or execute it"""
    pkg = os.path.join(ROOT, "psvo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "psvo_oracle" not in txt, f
