"""Exactly tied configuration costs: punctured symbols (all-zero LLR vectors) and quantised LLRs.

With continuous LLRs two T-EMS configurations of a syndrome never have exactly the same cost; with erasures or integer
LLRs they do all the time, and which of the tied paths is RECORDED changes the messages.  tie_mode = REFERENCE must
reproduce the path the reference's depth-first enumeration keeps (NBLDPC.cpp:1892-1944) -- on the fast T-EMS kernel too.  All values in these tests are
dyadic rationals, so fp32 and fp64 arithmetic are both exact and BOTH precisions must match the oracle decision for
decision.  EMS and QSPA have one kernel each; their tie rules (insertion-sort order, NBLDPC.cpp:1715-1746; first
maximum) are checked on the same inputs."""
import numpy as np
import pytest

from test_warp_paths_gpu import _random_code

pytestmark = pytest.mark.gpu


def _run(code, tmp_path, lch, method, p0, p1, offset, precision, max_iter=5, tie_mode=2):
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    path = tmp_path / ("tie_%d_%d.txt" % (code["q"], code["N"]))
    path.write_text(po.code_text(code))
    oc = po.OracleCode(code)
    want = [oc.decode(method, max_iter, p0, p1, 1.0, offset, l) for l in lch]
    dec = capi.Decoder(capi.Code.from_file(str(path)), method, max_iter, p0, p1, 1.0, offset, precision=precision,
                       max_batch=len(lch), tie_mode=tie_mode)
    r = dec.decode(lch, want_post=True)
    r["kernel"] = dec.kernel()
    return r, want


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("q,regular,p0,p1", [(16, False, 2, 2), (32, False, 2, 2), (64, True, 2, 2), (256, True, 2, 2),
                                             (64, True, 1, 2), (64, True, 3, 2), (32, False, 2, 1), (64, False, 3, 2),
                                             (256, False, 2, 2), (256, True, 1, 2), (256, False, 2, 1), (128, True, 2, 2), (128, False, 1, 2)])
def test_tems_integer_llrs_and_erasures(q, regular, p0, p1, precision, tmp_path):
    rng = np.random.default_rng(4000 + q + 10 * p0 + p1)
    code = _random_code(q, 8, rng, regular)
    F = 16
    lch = rng.integers(-6, 3, size=(F, code["N"], q - 1)).astype(np.float64)     # integer LLRs: ties everywhere
    lch[:, ::3, :] = 0.0                                                          # every third symbol erased
    r, want = _run(code, tmp_path, lch, 4, p0, p1, 0.5, precision)
    # the reference's tie order runs on the one-warp-per-row kernel (second pass over the handled entries) for q >= 32, n_r <= 2;
    # n_r = 3 (a third column can tie) and GF(16) on the group kernel
    assert r["kernel"] == ("tems_warp" if q >= 32 and p0 <= 2 else "group"), r["kernel"]
    assert np.array_equal(r["converged"], np.array([w[0] for w in want], np.int32))
    assert np.array_equal(r["iterations"], np.array([w[1] for w in want], np.int32))
    assert np.array_equal(r["sym"], np.stack([w[2] for w in want]))
    assert np.array_equal(r["post"], np.stack([w[3] for w in want]))              # exact arithmetic: identical posteriors


@pytest.mark.parametrize("precision", [0, 1])
def test_ems_integer_llrs_and_erasures(precision, tmp_path):
    rng = np.random.default_rng(4100)
    for q, regular in ((16, False), (64, True)):
        code = _random_code(q, 8, rng, regular)
        lch = rng.integers(-6, 3, size=(12, code["N"], q - 1)).astype(np.float64)
        lch[:, ::3, :] = 0.0
        r, want = _run(code, tmp_path, lch, 2, 8, 3, 0.5, precision)
        assert np.array_equal(r["iterations"], np.array([w[1] for w in want], np.int32))
        assert np.array_equal(r["sym"], np.stack([w[2] for w in want]))
        assert np.array_equal(r["post"], np.stack([w[3] for w in want]))


def test_punctured_frames_through_the_front_end(tmp_path):
    """BPSK frames with punctured symbols: demapper (zero LLR vectors) + T-EMS, fp64, against the oracle chain"""
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    rng = np.random.default_rng(4200)
    for q, regular in ((64, True), (256, True), (32, False)):
        code = _random_code(q, 8, rng, regular)
        p = q.bit_length() - 1
        N = code["N"]
        punct = sorted(rng.choice(N, size=N // 4, replace=False).tolist())
        path = tmp_path / ("pun_%d.txt" % q)
        path.write_text(po.code_text(code))
        F, sigma = 10, 0.55
        y = 1.0 + sigma * rng.standard_normal((F, (N - len(punct)) * p))
        oc = po.OracleCode(code)
        want = []
        for f in range(F):
            L = po.demap_bpsk(q, N, y[f], sigma, punct)
            assert not L[punct].any()
            want.append(oc.decode(4, 8, 2, 2, 1.0, 0.5, L))
        dec = capi.Decoder(capi.Code.from_file(str(path)), 4, 8, 2, 2, 1.0, 0.5, precision=capi.PREC_F64, max_batch=F)
        r = dec.demap_decode(2, y, sigma, punct=punct)
        assert np.array_equal(r["iterations"], np.array([w[1] for w in want], np.int32)), q
        assert np.array_equal(r["sym"], np.stack([w[2] for w in want])), q
        assert len(set(r["iterations"].tolist())) > 1


def test_64qam_front_end_with_punctured_symbols(code_files, tmp_path):
    """GF(64) + Gray 64-QAM: device demapper bit-exact with the oracle's, punctured symbols included,
    and T-EMS / EMS decisions identical to the oracle on the resulting LLRs (fp64)"""
    from conftest import constellation, get_code_dict, load_product_code
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    cs = constellation("GRAY_64QAM")
    cd = get_code_dict("BDS.576.288.GF.64")
    code = load_product_code(code_files, "BDS.576.288.GF.64")
    rng = np.random.default_rng(4300)
    N, F = cd["N"], 6
    punct = sorted(rng.choice(N, size=6, replace=False).tolist())
    sigma = 0.11
    sent = rng.integers(0, 64, size=(F, N))                       # not code words: the decoder is compared, not the code
    keep = [s for s in range(N) if s not in punct]
    y = cs[sent[:, keep]] + sigma * rng.standard_normal((F, len(keep), 2))
    oc = po.OracleCode(cd)
    L = np.stack([po.demap_qam(64, N, y[f], cs, sigma, punct) for f in range(F)])
    for method, p0, p1 in ((4, 2, 2), (2, 8, 3)):
        dec = capi.Decoder(code, method, 6, p0, p1, 1.0, 0.5, precision=capi.PREC_F64, max_batch=F)
        got = dec.demap_qam(y, cs, sigma, punct=punct)
        assert np.array_equal(got, L)
        want = [oc.decode(method, 6, p0, p1, 1.0, 0.5, L[f]) for f in range(F)]
        r = dec.demap_decode(64, y, sigma, constellation=cs, punct=punct)
        assert np.array_equal(r["iterations"], np.array([w[1] for w in want], np.int32))
        assert np.array_equal(r["sym"], np.stack([w[2] for w in want]))


@pytest.mark.parametrize("q,regular,p0,p1,factor", [(16, True, 3, 2, 1.25), (32, False, 2, 2, 0.8), (64, True, 2, 3, 1.25), (16, False, 1, 3, 1.0)])
def test_tems_bitexact_mode_on_inputs_no_scan_order_can_promise(q, regular, p0, p1, factor, tmp_path):
    """Quantised samples with a non-dyadic scaling factor and punctured symbols with continuous samples: the costs of
    tied configurations differ only by the rounding residue the reference's running sum carries at that point of its
    search.  Tie mode BITEXACT replays the search, so posteriors must equal the oracle's bit for bit (the oracle is
    pinned bit for bit to the compiled reference)."""
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    rng = np.random.default_rng(4300 + q + p0 + 7 * p1)
    code = _random_code(q, 8, rng, regular)
    p = q.bit_length() - 1
    N = code["N"]
    punct = [1, N // 2]
    F = 6
    y = 1.0 + 0.8 * rng.standard_normal((F, (N - len(punct)) * p))
    y[F // 2:] = np.round(y[F // 2:] * 2) / 2                                     # half of the frames quantised
    lch = np.stack([po.demap_bpsk(q, N, yy, 0.8, punct) for yy in y])
    path = tmp_path / "bx.txt"
    path.write_text(po.code_text(code))
    oc = po.OracleCode(code)
    want = [oc.decode(4, 7, p0, p1, factor, 0.25, l) for l in lch]
    dec = capi.Decoder(capi.Code.from_file(str(path)), 4, 7, p0, p1, factor, 0.25, precision=capi.PREC_F64, max_batch=F,
                       tie_mode=capi.TIES_BITEXACT)
    r = dec.decode(lch, want_post=True)
    assert np.array_equal(r["iterations"], np.array([w[1] for w in want], np.int32))
    assert np.array_equal(r["sym"], np.stack([w[2] for w in want]))
    assert np.array_equal(r["post"], np.stack([w[3] for w in want]))
