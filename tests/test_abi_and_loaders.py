"""CPU tests of the C-ABI library: it loads, exports every symbol include/nbldpc_b200.h declares, and its host-side
loaders (GF tables, code files, H_bit image) agree with the independent Python parsers.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REFERENCE, ROOT, bundle, get_code_dict, load_product_code
from nbldpc_b200 import capi
from oracle import pyoracle as po


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nbldpc_b200.h")).read()
    names = sorted(set(re.findall(r"\b(nbl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = ctypes.CDLL(capi.lib_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert capi.lib().nbl_abi_version() == 4


def test_gf_tables_match_oracle_and_reference_files():
    for q in (4, 8, 16, 32, 64, 128, 256):
        mul, inv = capi.gf_tables(q)
        m2, i2 = po.gf_tables(q)
        assert np.array_equal(mul, m2) and np.array_equal(inv, i2)
        path = os.path.join(REFERENCE, "Arith.Table.GF.%d.txt" % q)
        if os.path.exists(path):
            assert capi.gf_check_table_file(path, q) == 0
    with pytest.raises(capi.NblError):
        capi.gf_tables(16, 21)           # x^4+x^2+1 is not primitive


def test_code_file_loader_matches_python_parser(code_files):
    for name in bundle()["codes"]:
        ref = get_code_dict(name)
        code = load_product_code(code_files, name)
        assert (code.N, code.M, code.q, code.E) == (ref["N"], ref["M"], ref["q"], sum(ref["cdeg"]))
        assert code.max_dv == max(ref["vdeg"]) and code.max_dc == max(ref["cdeg"])
        g = code.graph()
        for r, links in enumerate(ref["chk"]):
            lo, hi = g["chk_ptr"][r], g["chk_ptr"][r + 1]
            assert list(zip(g["chk_var"][lo:hi], g["chk_coef"][lo:hi])) == links
            for d, (v, _) in enumerate(links):                      # ChkLinkDv / VarLinkDc cross indices
                dv = g["chk_dv"][lo + d]
                s = g["var_ptr"][v] + dv
                assert g["var_chk"][s] == r and g["var_dc"][s] == d
        for v, links in enumerate(ref["var"]):
            lo, hi = g["var_ptr"][v], g["var_ptr"][v + 1]
            assert list(zip(g["var_chk"][lo:hi], g["var_coef"][lo:hi])) == links


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
def test_loader_reads_the_reference_files_directly():
    for name in bundle()["codes"]:
        code = capi.Code.from_file(os.path.join(REFERENCE, name + ".txt"))
        ref = get_code_dict(name)
        assert (code.N, code.M, code.q) == (ref["N"], ref["M"], ref["q"])
    code = capi.Code.from_hbit(os.path.join(REFERENCE, "H_bit_GF256.txt"), 256)
    assert (code.N, code.M, code.E) == (16, 8, 32)
    img = capi.Code.from_hbit(os.path.join(REFERENCE, "H_bit_GF64.txt"), 64)       # == image of BDS.576.288.GF.64
    bds = get_code_dict("BDS.576.288.GF.64")
    g = img.graph()
    assert (img.N, img.M) == (96, 48)
    for r, links in enumerate(bds["chk"]):
        lo, hi = g["chk_ptr"][r], g["chk_ptr"][r + 1]
        assert list(zip(g["chk_var"][lo:hi], g["chk_coef"][lo:hi])) == links


def test_hbit_loader(code_files):
    code = load_product_code(code_files, "H_bit_GF256")
    ref = get_code_dict("H_bit_GF256")
    g = code.graph()
    assert (code.N, code.M, code.E, code.max_dv, code.max_dc) == (16, 8, 32, 2, 4)
    for r, links in enumerate(ref["chk"]):
        lo, hi = g["chk_ptr"][r], g["chk_ptr"][r + 1]
        assert list(zip(g["chk_var"][lo:hi], g["chk_coef"][lo:hi])) == links


def test_loader_error_paths(tmp_path):
    with pytest.raises(capi.NblError):
        capi.Code.from_file(str(tmp_path / "missing.txt"))
    bad = tmp_path / "bad.txt"
    bad.write_text("4 2 16\n2 4\n2 2 2 2\n4 4\n1 1 2 1\n")                     # truncated
    with pytest.raises(capi.NblError):
        capi.Code.from_file(str(bad))
    txt = po.code_text(get_code_dict("divsalar.UNBLDPC.128.64.GF.16")).split("\n")
    txt[4] = txt[4].replace(txt[4].split()[1], "9", 1) if txt[4].split()[1] != "9" else txt[4].replace("9", "8", 1)
    inconsistent = tmp_path / "inconsistent.txt"
    inconsistent.write_text("\n".join(txt))
    with pytest.raises(capi.NblError):
        capi.Code.from_file(str(inconsistent))                                  # coefficient differs between the two lists
    with pytest.raises(capi.NblError):
        capi.Code.from_checks(512, 4, [[(0, 1), (1, 1)]])                        # q out of range


def test_puncture_rule():
    code = capi.Code.from_checks(16, 4, [[(0, 1), (1, 2), (2, 3)], [(1, 1), (2, 1), (3, 5)]])
    assert list(code.puncture_set(2)) == [1, 2] and list(code.puncture_set(1)) == [0, 3] and len(code.puncture_set(0)) == 0


def test_compute_entry_points_fail_loudly_without_a_gpu():
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    code = capi.Code.from_checks(16, 4, [[(0, 1), (1, 2), (2, 3)], [(1, 1), (2, 1), (3, 5)]])
    with pytest.raises(capi.NblError, match="no CPU fallback"):
        capi.Decoder(code, capi.METHOD_TEMS, 10, 2, 2, 1.0, 0.5)


def test_gf_tables_reject_polynomials_of_the_wrong_degree():
    """a polynomial without bit q, or with bits above it, must be refused (it would index past the log tables)"""
    for q, poly in ((64, 19), (16, 19 | 64), (16, 3), (256, 285 << 1), (16, -19)):
        with pytest.raises(capi.NblError):
            capi.gf_tables(q, poly)
    capi.gf_tables(16, 25)                  # x^4 + x^3 + 1: the other primitive polynomial of GF(16) is accepted


def test_code_create_argument_checks():
    ok = [[(0, 1), (1, 2)], [(1, 3), (2, 1)]]
    capi.Code.from_checks(16, 3, ok)
    with pytest.raises(capi.NblError):      # the same (row, variable) edge twice
        capi.Code.from_checks(16, 3, [[(0, 1), (0, 2)], [(1, 3), (2, 1)]])
    with pytest.raises(capi.NblError):      # variable index out of range
        capi.Code.from_checks(16, 3, [[(0, 1), (3, 2)], [(1, 3), (2, 1)]])
    with pytest.raises(capi.NblError):      # coefficient out of range
        capi.Code.from_checks(16, 3, [[(0, 1), (1, 16)], [(1, 3), (2, 1)]])
    L = capi.lib()
    h = ctypes.c_void_p()
    cptr = np.array([0, 2, 1], np.int32); cvar = np.array([0, 1], np.int32); ccoef = np.array([1, 2], np.int32)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    assert L.nbl_code_create(ctypes.byref(h), 16, 0, 3, 2, ip(cptr), ip(cvar), ip(ccoef), None, None, None) < 0     # chk_ptr decreases
    vptr = np.array([0, 1, 2, 2], np.int32)
    cptr = np.array([0, 1, 2], np.int32)
    assert L.nbl_code_create(ctypes.byref(h), 16, 0, 3, 2, ip(cptr), ip(cvar), ip(ccoef), ip(vptr), None, None) < 0  # var_ptr without lists


def test_no_gpu_no_fallback_for_the_multi_gpu_entry_points():
    """the communicator refuses to exist without its devices (and never falls back to a host reduction)"""
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.NblError):
        capi.Comm(1)
    with pytest.raises(capi.NblError):
        capi.Comm.from_rank(1, 0, None, 0)


def test_nothing_in_the_library_reads_the_environment():
    """kernel selection is a field of nbl_decoder_config; a stray environment variable must not move a decode to another kernel"""
    src = os.path.join(ROOT, "nbldpc_b200")
    for sub in ("csrc", "host"):
        for f in os.listdir(os.path.join(src, sub)):
            if f.endswith((".cu", ".cuh", ".cpp", ".h")):
                assert "getenv" not in open(os.path.join(src, sub, f)).read(), f
