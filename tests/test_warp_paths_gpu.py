"""GPU parity tests for every instantiation of the T-EMS warp kernel (nbldpc_b200/csrc/tems_warp.cuh).

The golden cases only reach the (check degree 4, variable degree 2) specialisation at q = 64 and q = 256.  Here the same
kernel is driven through its degree-generic form, through all four vector widths (q = 32, 64, 128, 256) and through
the n_r = 1 / n_c = 1 / factor != 1 branches, on random regular and irregular codes, against the C oracle (which is
pinned bit for bit to the compiled reference on the golden cases).  fp64 parity mode: return code, iteration count and
every hard decision must be identical; posteriors to 1e-8."""
import os

import numpy as np
import pytest

from conftest import golden_lch, load_golden, load_product_code

pytestmark = pytest.mark.gpu


def _random_code(q, M, rng, regular):
    """check-side edge lists of a random code; regular: dc = 4, dv = 2 (N = 2M); else dc in 2..4, dv in 1..5"""
    from oracle import pyoracle as po
    N = 2 * M if regular else (3 * M) // 2
    while True:
        if regular:
            stubs = rng.permutation(np.repeat(np.arange(N), 2))
            rows = [list(stubs[4 * r:4 * r + 4]) for r in range(M)]
        else:
            deg = np.zeros(N, int)
            rows = []
            for r in range(M):
                dc = int(rng.integers(2, 5))
                order = sorted(range(N), key=lambda v: (deg[v], rng.random()))     # lowest degree first: dv stays small
                pick = order[:dc]
                for v in pick:
                    deg[v] += 1
                rows.append(pick)
            if deg.min() < 1 or deg.max() > 5:
                continue
        if all(len(set(r)) == len(r) for r in rows):
            break
    chk = [[(int(v), int(rng.integers(1, q))) for v in sorted(r)] for r in rows]
    return po.code_from_checks(N, M, q, chk)


def _frames(code, F, sigma, rng):
    """all-zero codeword over BPSK/AWGN -> channel LLRs [F][N][q-1] through the oracle's demapper"""
    from oracle import pyoracle as po
    p = code["q"].bit_length() - 1
    return np.stack([po.demap_bpsk(code["q"], code["N"], 1.0 + sigma * rng.standard_normal(code["N"] * p), sigma) for _ in range(F)])


def _check_against_oracle(code, tmp_path, lch, p0, p1, factor, offset, kernel=0, max_iter=6, tie_mode=1):
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    path = tmp_path / ("code_%d_%d.txt" % (code["q"], code["N"]))
    path.write_text(po.code_text(code))
    oc = po.OracleCode(code)
    want = [oc.decode(4, max_iter, p0, p1, factor, offset, l) for l in lch]
    dec = capi.Decoder(capi.Code.from_file(str(path)), 4, max_iter, p0, p1, factor, offset, precision=capi.PREC_F64,
                       max_batch=len(lch), tie_mode=tie_mode, kernel=kernel)               # 1: the warp kernel also in fp64
    r = dec.decode(lch, want_post=True)
    assert np.array_equal(r["converged"], np.array([w[0] for w in want], np.int32))
    assert np.array_equal(r["iterations"], np.array([w[1] for w in want], np.int32))
    assert np.array_equal(r["sym"], np.stack([w[2] for w in want]))
    ref = np.stack([w[3] for w in want])
    err = np.abs(r["post"] - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 1e-8, err.max()
    return r


@pytest.mark.parametrize("q", [32, 64, 128, 256])
@pytest.mark.parametrize("regular", [True, False])
def test_random_codes_all_vector_widths(q, regular, tmp_path):
    rng = np.random.default_rng(1000 + q + regular)
    code = _random_code(q, 8, rng, regular)
    p = q.bit_length() - 1
    sigma = 0.62 + 0.02 * p                                       # noisy enough for several iterations and some failures
    lch = _frames(code, 10, sigma, rng)
    r = _check_against_oracle(code, tmp_path, lch, 2, 2, 1.0, 0.5)
    assert len(set(r["iterations"].tolist())) > 1                   # the cases are not all trivial


@pytest.mark.parametrize("p0,p1,factor,offset", [(1, 2, 1.0, 0.5), (3, 2, 1.0, 0.25), (2, 1, 1.0, 0.5), (2, 2, 1.5, 0.0), (2, 2, 0.8, 0.3)])
def test_parameter_branches(p0, p1, factor, offset, tmp_path):
    """n_r = 1 (second column forbidden), n_r = 3, n_c = 1 (no pair search), factor != 1 (division), offset = 0"""
    rng = np.random.default_rng(77)
    for q, regular in ((64, True), (32, False)):
        code = _random_code(q, 8, rng, regular)
        lch = _frames(code, 6, 0.75, rng)
        _check_against_oracle(code, tmp_path, lch, p0, p1, factor, offset)


@pytest.mark.parametrize("kernel", [2, 1])      # NBL_KERNEL_WARP_GENERIC, NBL_KERNEL_GROUP
@pytest.mark.parametrize("name", ["cfg3_tems_gf256", "cfg5_tems_hbit256"])
def test_golden_through_the_other_kernels(name, kernel, code_files):
    """the golden T-EMS frames through the degree-generic warp kernel and through the group kernel"""
    from nbldpc_b200 import capi
    z, meta = load_golden(name)
    code = load_product_code(code_files, meta["code"])
    F = len(z["ret"])
    dec = capi.Decoder(code, meta["method"], meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"],
                       precision=capi.PREC_F64, max_batch=F, tie_mode=1, kernel=kernel)
    r = dec.decode(golden_lch(z, meta))
    assert np.array_equal(r["converged"], z["ret"].astype(np.int32))
    assert np.array_equal(r["iterations"], z["iters"].astype(np.int32))
    assert np.array_equal(r["sym"], z["dec_sym"].astype(np.int32))


def test_fp32_regular_and_generic_kernels_agree(code_files):
    """fp32 fast mode: both warp-kernel specialisations perform the same operations in the same order"""
    from nbldpc_b200 import capi
    z, meta = load_golden("cfg3_tems_gf256")
    code = load_product_code(code_files, meta["code"])
    F = len(z["ret"])
    lch = golden_lch(z, meta)
    outs = []
    for kernel in (0, 2):
        dec = capi.Decoder(code, meta["method"], meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"],
                           precision=capi.PREC_F32, max_batch=F, kernel=kernel)
        outs.append(dec.decode(lch, want_post=True))
    for k in ("sym", "converged", "iterations", "post"):
        assert np.array_equal(outs[0][k], outs[1][k]), k


def _check_ems(code, tmp_path, lch, nm, nc, factor, offset, kernel=0, max_iter=6):
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    path = tmp_path / ("ems_%d_%d.txt" % (code["q"], code["N"]))
    path.write_text(po.code_text(code))
    oc = po.OracleCode(code)
    want = [oc.decode(2, max_iter, nm, nc, factor, offset, l) for l in lch]
    dec = capi.Decoder(capi.Code.from_file(str(path)), 2, max_iter, nm, nc, factor, offset, precision=capi.PREC_F64, max_batch=len(lch), kernel=kernel)
    r = dec.decode(lch, want_post=True)
    assert np.array_equal(r["converged"], np.array([w[0] for w in want], np.int32))
    assert np.array_equal(r["iterations"], np.array([w[1] for w in want], np.int32))
    assert np.array_equal(r["sym"], np.stack([w[2] for w in want]))
    ref = np.stack([w[3] for w in want])
    err = np.abs(r["post"] - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 1e-8, err.max()
    return r


@pytest.mark.parametrize("q", [32, 64, 128, 256])
@pytest.mark.parametrize("regular", [True, False])
def test_ems_warp_kernel_random_codes(q, regular, tmp_path):
    """the EMS warp kernel (ems_warp.cuh): all vector widths, both degree specialisations, n_m / n_c / factor branches"""
    rng = np.random.default_rng(2000 + q + regular)
    code = _random_code(q, 8, rng, regular)
    p = q.bit_length() - 1
    lch = _frames(code, 8, 0.62 + 0.02 * p, rng)
    r = _check_ems(code, tmp_path, lch, 8, 3, 1.0, 0.5)
    assert len(set(r["iterations"].tolist())) > 1
    _check_ems(code, tmp_path, lch[:4], 4, 2, 1.0, 0.25)
    _check_ems(code, tmp_path, lch[:4], 8, 1, 1.3, 0.0)
    _check_ems(code, tmp_path, lch[:4], 16, 2, 1.0, 0.5)
    _check_ems(code, tmp_path, lch[:4], 6, 3, 1.0, 0.5)                              # n_m not a power of two: group kernel
    _check_ems(code, tmp_path, lch[:4], 8, 3, 1.0, 0.5, kernel=2)


@pytest.mark.parametrize("q,nm,nc", [(64, 8, 3), (64, 16, 3), (64, 2, 3), (64, 8, 2), (32, 8, 3), (128, 4, 3)])
def test_ems_fp32_two_stage_enumeration_equals_the_group_kernel(q, nm, nc, tmp_path):
    """fp32: the EMS warp kernel -- conf(n_m, n_c) in two stages when n_c does not bind, with the two-copy table, FADD2 and FMNMX3
    for GF(64); direct enumeration otherwise -- against the group kernel's dynamic programme on the same frames: both form
    every sum left to right and keep a value, not a path, so the outputs are identical bit for bit"""
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    rng = np.random.default_rng(5200 + q + nm + nc)
    code = _random_code(q, 8, rng, True)
    lch = _frames(code, 12, 0.62 + 0.02 * (q.bit_length() - 1), rng)
    path = tmp_path / ("ems32_%d.txt" % q)
    path.write_text(po.code_text(code))
    outs = []
    for kernel in (0, 1):                                                     # AUTO (ems_warp), GROUP
        dec = capi.Decoder(capi.Code.from_file(str(path)), 2, 6, nm, nc, 1.0, 0.5, precision=capi.PREC_F32, max_batch=len(lch), kernel=kernel)
        assert dec.kernel() == ("ems_warp" if kernel == 0 else "group")
        outs.append(dec.decode(lch, want_post=True))
    for k in ("sym", "converged", "iterations", "post"):
        assert np.array_equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("q", [16, 32, 64, 128])
@pytest.mark.parametrize("regular", [True, False])
def test_qspa_warp_kernel_random_codes(q, regular, tmp_path):
    """the QSPA warp kernel (qspa_warp.cuh; q = 16: the half-warp kernel, qspa_half.cuh) against the oracle's log-domain box-plus: decisions and iteration counts
    identical, posteriors to 1e-6 on the part above the transform's noise floor"""
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    rng = np.random.default_rng(3000 + q + regular)
    code = _random_code(q, 6, rng, regular)
    p = q.bit_length() - 1
    lch = _frames(code, 5 if q < 128 else 3, 0.60 + 0.02 * p, rng)
    path = tmp_path / ("qspa_%d.txt" % q)
    path.write_text(po.code_text(code))
    oc = po.OracleCode(code)
    want = [oc.decode(1, 5, 0, 0, 1.0, 0.0, l) for l in lch]
    dec = capi.Decoder(capi.Code.from_file(str(path)), 1, 5, precision=capi.PREC_F64, max_batch=len(lch))
    r = dec.decode(lch, want_post=True)
    assert np.array_equal(r["converged"], np.array([w[0] for w in want], np.int32))
    assert np.array_equal(r["iterations"], np.array([w[1] for w in want], np.int32))
    assert np.array_equal(r["sym"], np.stack([w[2] for w in want]))
    ref = np.stack([w[3] for w in want])
    sel = ref > -25.0
    err = np.abs(r["post"] - ref)[sel] / np.maximum(np.abs(ref[sel]), 1.0)
    assert err.max() < 1e-6, err.max()


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("name", ["cfg1_qspa_gf16", "cfg1_qspa_gf16_allzero"])
def test_qspa_half_warp_and_group_kernels_are_bit_identical(name, precision, code_files):
    """GF(16): the half-warp kernel (registers and shuffles only) and the group kernel agree bit for bit, posteriors included"""
    from nbldpc_b200 import capi
    z, meta = load_golden(name)
    code = load_product_code(code_files, meta["code"])
    lch = golden_lch(z, meta)
    outs = []
    for kernel in (0, 1):
        dec = capi.Decoder(code, 1, meta["max_iter"], precision=precision, max_batch=len(lch), kernel=kernel)
        outs.append(dec.decode(lch, want_post=True))
    for k in ("sym", "converged", "iterations", "post"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    if precision == capi.PREC_F64:
        assert np.array_equal(outs[0]["sym"], z["dec_sym"].astype(np.int32))


@pytest.mark.parametrize("precision", [0, 1])
def test_qspa_warp_and_group_kernels_are_bit_identical(precision, code_files):
    """same butterflies in the same order: the warp form and the group form of the Walsh-Hadamard check node agree bit for bit"""
    from conftest import constellation
    from nbldpc_b200 import capi
    z, meta = load_golden("cfg4_qspa_gf256_qam")
    code = load_product_code(code_files, meta["code"])
    lch = golden_lch(z, meta)
    outs = []
    for kernel in (0, 1):
        dec = capi.Decoder(code, 1, meta["max_iter"], precision=precision, max_batch=len(lch), kernel=kernel)
        outs.append(dec.decode(lch, want_post=True))
    for k in ("sym", "converged", "iterations", "post"):
        assert np.array_equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("name", ["f2_bstems_gf256", "f2_bstems_gf64_nm9", "f2_bstems_gf16"])
def test_bstems_warp_and_group_kernels_agree(name, precision, code_files):
    """BS-TEMS on the warp kernel (tems_warp.cuh, BS = 1: basic-set rounds + one lane walking the subsets) and on the group
    kernel (cn_bstems): same decisions, iteration counts and posteriors (both form every sum in the reference's order)"""
    from nbldpc_b200 import capi
    z, meta = load_golden(name)
    code = load_product_code(code_files, meta["code"])
    lch = golden_lch(z, meta)
    outs = []
    for kernel in (0, 1):
        dec = capi.Decoder(code, meta["method"], meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"],
                           precision=precision, max_batch=len(lch), kernel=kernel)
        outs.append(dec.decode(lch, want_post=True))
    for k in ("sym", "converged", "iterations"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    assert np.abs(outs[0]["post"] - outs[1]["post"]).max() < (1e-9 if precision == capi.PREC_F64 else 1e-3)
    if precision == capi.PREC_F64:
        assert np.array_equal(outs[0]["sym"], z["dec_sym"].astype(np.int32))
        assert np.array_equal(outs[0]["iterations"], z["iters"].astype(np.int32))


@pytest.mark.parametrize("regular", [True, False])
@pytest.mark.parametrize("p0,p1,factor,offset", [(2, 2, 1.0, 0.5), (1, 2, 1.0, 0.5), (3, 2, 0.8, 0.25), (2, 1, 1.0, 0.5), (2, 2, 1.5, 0.0)])
def test_tems_gf16_kernel_random_codes(regular, p0, p1, factor, offset, tmp_path):
    """T-EMS over GF(16) on the four-lanes-per-row kernel (tems_half.cuh; tie mode FAST) against the oracle: random regular
    and irregular codes (check degrees up to 5-6), n_r = 1 / 2 / 3, n_c = 1 / 2, scaling factor, zero offset"""
    rng = np.random.default_rng(1600 + regular + 10 * p0 + p1)
    code = _random_code(16, 12, rng, regular)
    lch = _frames(code, 16, 0.70, rng)
    r = _check_against_oracle(code, tmp_path, lch, p0, p1, factor, offset)
    assert len(set(r["iterations"].tolist())) > 1


def test_tems_gf16_kernel_on_the_reference_code(code_files):
    """the (128,64) GF(16) code of the reference (d_c = 5, d_v = 2 and 3), frames of the QSPA fixture decoded with T-EMS:
    the GF(16) kernel and the group kernel give the same decisions and iteration counts in fp64 and agree in fp32"""
    from nbldpc_b200 import capi
    z, meta = load_golden("cfg1_qspa_gf16")
    code = load_product_code(code_files, meta["code"])
    lch = golden_lch(z, meta)
    for precision, tol in ((capi.PREC_F64, 1e-8), (capi.PREC_F32, 1e-3)):
        outs = []
        for kernel in (0, 1):
            dec = capi.Decoder(code, 4, 10, 2, 2, 1.0, 0.5, precision=precision, max_batch=len(lch), tie_mode=1, kernel=kernel)
            outs.append(dec.decode(lch, want_post=True))
        same = (outs[0]["sym"] == outs[1]["sym"]).all(axis=1) & (outs[0]["iterations"] == outs[1]["iterations"])
        assert same.all() if precision == capi.PREC_F64 else same.mean() > 0.9
        d = np.abs(outs[0]["post"] - outs[1]["post"])[same]
        assert (d / np.maximum(np.abs(outs[1]["post"][same]), 1.0)).max() < tol


def test_baseline_configurations_run_on_their_fast_kernels(code_files):
    """every BASELINE configuration (and the other decoders on the reference's codes) selects the kernel built for it; the
    parity modes and the parameters outside a fast kernel's range fall back to the group kernel"""
    from nbldpc_b200 import capi
    G16, G64, G256 = "divsalar.UNBLDPC.128.64.GF.16", "BDS.576.288.GF.64", "divsalar.UNBLDPC.512.256.GF.256"
    want = [(G16, 1, 0, 0, 0.0, "qspa_gf16"), (G64, 2, 8, 3, 0.5, "ems_warp"), (G256, 4, 2, 2, 0.5, "tems_warp"),
            ("divsalar.CNBLDPC.512.256.GF.256", 1, 0, 0, 0.0, "qspa_warp"), ("H_bit_GF256", 4, 2, 2, 0.5, "tems_warp"),
            (G256, 7, 4, 3, 0.2, "bstems_warp"), (G16, 4, 2, 2, 0.5, "tems_gf16"), (G64, 4, 2, 2, 0.5, "tems_warp"),
            (G256, 4, 2, 3, 0.5, "group"), (G16, 2, 8, 3, 0.5, "ems_gf16"), (G16, 7, 4, 3, 0.2, "bstems_gf16"), (G64, 2, 6, 3, 0.5, "group")]
    for name, method, p0, p1, off, kern in want:
        code = load_product_code(code_files, name)
        dec = capi.Decoder(code, method, 10, p0, p1, 1.0, off, precision=capi.PREC_F32, max_batch=4)
        assert dec.kernel() == kern, (name, method, p0, p1, dec.kernel())
    code = load_product_code(code_files, G256)
    # the fp64 default (reference tie order) runs on the fast kernel too; the bit-exact replay and n_r = 3 with the reference order do not
    assert capi.Decoder(code, 4, 10, 2, 2, 1.0, 0.5, precision=capi.PREC_F64, max_batch=4).kernel() == "tems_warp"
    assert capi.Decoder(code, 4, 10, 2, 2, 1.0, 0.5, precision=capi.PREC_F64, max_batch=4, tie_mode=1).kernel() == "tems_warp"
    assert capi.Decoder(code, 4, 10, 2, 2, 1.0, 0.5, precision=capi.PREC_F32, max_batch=4, tie_mode=2).kernel() == "tems_warp"
    assert capi.Decoder(code, 4, 10, 2, 2, 1.0, 0.5, precision=capi.PREC_F64, max_batch=4, tie_mode=3).kernel() == "group"
    assert capi.Decoder(code, 4, 10, 3, 2, 1.0, 0.5, precision=capi.PREC_F64, max_batch=4, tie_mode=2).kernel() == "group"


@pytest.mark.parametrize("regular", [True, False])
@pytest.mark.parametrize("nm,nc,factor,offset", [(8, 3, 1.0, 0.5), (6, 2, 1.0, 0.2), (16, 1, 1.25, 0.0), (4, 3, 1.0, 0.5)])
def test_ems_gf16_kernel_random_codes(regular, nm, nc, factor, offset, tmp_path):
    """EMS over GF(16) on the four-lanes-per-row kernel (ems_half.cuh) against the oracle: random regular and irregular
    codes, n_m a power of two or not, n_c = 1..3"""
    rng = np.random.default_rng(1700 + regular + nm + nc)
    code = _random_code(16, 12, rng, regular)
    lch = _frames(code, 12, 0.70, rng)
    _check_ems(code, tmp_path, lch, nm, nc, factor, offset)


def test_ems_gf16_kernel_on_the_reference_code(code_files):
    """the golden EMS frames on the (128,64) GF(16) code (d_c = 5): reference decisions in fp64 on the GF(16) kernel, and
    the group kernel gives the same posteriors"""
    from nbldpc_b200 import capi
    z, meta = load_golden("x_ems_gf16_dc5")
    code = load_product_code(code_files, meta["code"])
    lch = golden_lch(z, meta)
    outs = []
    for kernel in (0, 1):
        dec = capi.Decoder(code, 2, meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"], precision=capi.PREC_F64,
                           max_batch=len(lch), kernel=kernel)
        outs.append((dec.kernel(), dec.decode(lch, want_post=True)))
    assert [k for k, _ in outs] == ["ems_gf16", "group"]
    for k in ("sym", "converged", "iterations", "post"):
        assert np.array_equal(outs[0][1][k], outs[1][1][k]), k
    assert np.array_equal(outs[0][1]["sym"], z["dec_sym"].astype(np.int32))
    assert np.array_equal(outs[0][1]["iterations"], z["iters"].astype(np.int32))


def test_ems_running_sum_residue_case(tmp_path):
    """The one EMS difference the randomised differential test has found (tools/fuzz_parity.py, seed 51, case 75: GF(128),
    irregular code, n_m = 4, n_c = 1): the reference forms all sums of a check row in one running sum (+= going down, -= coming
    back, NBLDPC.cpp:1746-1786), whose rounding residue decided between two sums that are equal in exact arithmetic; from
    iteration 5 on one variable's posterior differs.  Both CUDA kernels form every sum afresh: they agree with each other and
    with the oracle run in that arithmetic (an analysis switch of the oracle) bit for bit, and with the reference's
    arithmetic up to iteration 4."""
    import importlib.util, os
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(root, "tools", "fuzz_parity.py"))
    fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
    d = fz.gen_case(75, 51)
    code, q = d["code"], d["code"]["q"]
    assert (q, d["method"], d["p0"], d["p1"]) == (128, 2, 4, 1)
    L = np.stack([po.demap_bpsk(q, code["N"], yy, d["sigma"], d["punct"]) for yy in d["y"]])
    path = tmp_path / "residue.txt"
    path.write_text(po.code_text(code))
    oc = po.OracleCode(code)
    outs = []
    for kernel in (0, 1):
        dec = capi.Decoder(capi.Code.from_file(str(path)), 2, d["max_iter"], 4, 1, d["factor"], d["offset"], precision=capi.PREC_F64, max_batch=len(L), kernel=kernel)
        outs.append(dec.decode(L, want_post=True))
    assert all(np.array_equal(outs[0][k], outs[1][k]) for k in ("sym", "converged", "iterations", "post"))
    ref = np.stack([oc.decode(2, d["max_iter"], 4, 1, d["factor"], d["offset"], l)[3] for l in L])
    po.set_ems_fresh_sums(1)
    try:
        fresh = [oc.decode(2, d["max_iter"], 4, 1, d["factor"], d["offset"], l) for l in L]
    finally:
        po.set_ems_fresh_sums(0)
    assert np.abs(outs[0]["post"] - ref).max() > 1.0                              # the reference's residue took the other branch ...
    assert np.array_equal(outs[0]["sym"], np.stack([w[2] for w in fresh]))        # ... the same arithmetic without it does not
    assert np.array_equal(outs[0]["iterations"], np.array([w[1] for w in fresh], np.int32))
    assert np.abs(outs[0]["post"] - np.stack([w[3] for w in fresh])).max() < 1e-9
    ref4 = np.stack([oc.decode(2, 4, 4, 1, d["factor"], d["offset"], l)[3] for l in L])
    dec4 = capi.Decoder(capi.Code.from_file(str(path)), 2, 4, 4, 1, d["factor"], d["offset"], precision=capi.PREC_F64, max_batch=len(L))
    assert np.abs(dec4.decode(L, want_post=True)["post"] - ref4).max() < 1e-9
