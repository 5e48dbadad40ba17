"""bit-level OSD post-processing of failed frames (reference OSD.h).

CPU: the oracle's restatement against the fixtures dumped from the compiled reference (decisions of failed frames after
OSD, orders 1-3, CRC-aided rows, forgetting factor, OSD_flag, stand-alone method 6).  GPU: the CUDA path through the
C ABI against the same fixtures and against the oracle on random inputs."""
import numpy as np
import pytest

from conftest import get_code_dict, golden_lch, golden_osd_names, load_golden, load_product_code


def _oracle(meta):
    from oracle import pyoracle as po
    oc = po.OracleCode(get_code_dict(meta["code"]))
    return oc, po.OracleOsd(oc, meta["crc"], meta["crc_correct"])


@pytest.mark.parametrize("name", golden_osd_names())
def test_oracle_osd_matches_reference(name):
    z, meta = load_golden(name)
    if meta["q"] == 256 and meta["N"] > 16:
        pytest.skip("minutes on a CPU; pinned when the fixture was made (tests/golden/make_golden.py)")
    oc, osd = _oracle(meta)
    lch = golden_lch(z, meta)
    for f in range(len(z["ret"])):
        if meta["method"] == 6:
            assert np.array_equal(osd.run(meta["osd_order"], lch[f]), z["dec_sym"][f])
            continue
        r, it, out, _ = oc.decode(meta["method"], meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"], lch[f],
                                  osd, meta["osd_order"], meta["osd_factor"], meta["osd_flag"])
        assert (r, it) == (z["ret"][f], z["iters"][f])
        assert np.array_equal(out, z["dec_sym"][f]), f


def test_osd_improves_on_the_iterative_decoder():
    """sanity of the fixtures themselves: OSD touched exactly the failed frames and turned some of them into the
    transmitted code word (26 -> 19 frame errors at this SNR)"""
    z, meta = load_golden("f3_osd2_tems_gf16")
    oc, _ = _oracle(meta)
    lch = golden_lch(z, meta)
    plain = np.stack([oc.decode(meta["method"], meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"], l)[2] for l in lch])
    failed = z["ret"] == 0
    assert (plain[failed] != z["dec_sym"][failed]).any()
    assert np.array_equal(plain[~failed], z["dec_sym"][~failed])
    assert (z["dec_sym"] != z["tx_code_sym"]).any(axis=1).sum() < (plain != z["tx_code_sym"]).any(axis=1).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [1, 0])
@pytest.mark.parametrize("name", golden_osd_names())
def test_gpu_osd_matches_reference(name, precision, code_files):
    from nbldpc_b200 import capi
    z, meta = load_golden(name)
    code = load_product_code(code_files, meta["code"])
    F = len(z["ret"])
    lch = golden_lch(z, meta)
    method = meta["method"] if meta["method"] != 6 else 1
    dec = capi.Decoder(code, method, meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"],
                       precision=precision, max_batch=F)
    dec.enable_osd(meta["osd_order"], meta["osd_factor"], meta["osd_flag"], meta["crc"], meta["crc_correct"])
    if meta["method"] == 6:
        got = dec.osd(lch)
        same = (got == z["dec_sym"]).all(axis=1)
    else:
        r = dec.decode(lch)
        assert np.array_equal(r["converged"], z["ret"].astype(np.int32)) or precision == capi.PREC_F32
        same = (r["sym"] == z["dec_sym"]).all(axis=1) & (r["converged"] == z["ret"])
    if precision == capi.PREC_F64:
        assert same.all(), np.flatnonzero(~same)
    else:
        assert same.mean() >= 0.75, same          # fp32 messages: a near-tie in the ranking or a distance may fall the other way


@pytest.mark.gpu
@pytest.mark.parametrize("q,order,crc,rate,alpha,flag", [(16, 2, 8, 0, 1.0, 0), (32, 1, 8, 8, 0.5, 0), (64, 2, 16, 5, 1.0, 1), (16, 3, 8, 3, 0.7, 0)])
def test_gpu_osd_random_codes_against_oracle(q, order, crc, rate, alpha, flag, tmp_path):
    """random (some rank-deficient) codes: exercises the column rotation of the elimination and the candidate order"""
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    from test_warp_paths_gpu import _random_code
    rng = np.random.default_rng(5000 + q + order)
    code = _random_code(q, 10, rng, q != 32)
    p = q.bit_length() - 1
    N = code["N"]
    F = 12
    y = 1.0 + 1.0 * rng.standard_normal((F, N * p))
    lch = np.stack([po.demap_bpsk(q, N, yy, 1.0) for yy in y])
    path = tmp_path / "osd.txt"
    path.write_text(po.code_text(code))
    oc = po.OracleCode(code)
    osd = po.OracleOsd(oc, crc, rate)
    want = [oc.decode(4, 3, 2, 2, 1.0, 0.5, l, osd, order, alpha, flag) for l in lch]
    dec = capi.Decoder(capi.Code.from_file(str(path)), 4, 3, 2, 2, 1.0, 0.5, precision=capi.PREC_F64, max_batch=F)
    dec.enable_osd(order, alpha, flag, crc, rate)
    r = dec.decode(lch)
    assert np.array_equal(r["converged"], np.array([w[0] for w in want], np.int32))
    assert (r["converged"] == 0).any()
    assert np.array_equal(r["sym"], np.stack([w[2] for w in want]))
    assert np.array_equal(dec.osd(lch), np.stack([osd.run(order, l) for l in lch]))


@pytest.mark.gpu
def test_device_monte_carlo_with_osd(code_files):
    """nbl_simulate with OSD enabled: the same frames (same seed) lose frame errors, never gain one on a frame that
    converged; the CRC-row variant is refused unless the device-drawn messages carry that CRC (crc_len)"""
    from nbldpc_b200 import capi
    code = load_product_code(code_files, "divsalar.UNBLDPC.128.64.GF.16")
    frames = 20000
    dec = capi.Decoder(code, capi.METHOD_TEMS, 5, 2, 2, 1.0, 0.5, precision=capi.PREC_F32, max_batch=frames)
    plain = dec.simulate(2.0, frames, seed=3, random_msg=1)
    dec.enable_osd(1, 1.0, 0, 8, 0)
    osd = dec.simulate(2.0, frames, seed=3, random_msg=1)
    assert int(plain[0]) == int(osd[0]) == frames
    assert int(osd[4]) == int(plain[4])                     # converged frames: untouched
    assert 0 < int(osd[1]) < int(plain[1]), (plain, osd)    # frame errors: some failed frames are now the transmitted code word
    dec.enable_osd(1, 1.0, 0, 8, 8)
    with pytest.raises(Exception, match="same CRC"):
        dec.simulate(2.0, 1000, seed=3, random_msg=1)
    crc_rows = dec.simulate(2.0, frames, seed=3, random_msg=1, crc_len=8)     # messages now carry the CRC the rows check
    dec.enable_osd(-1)
    crc_plain = dec.simulate(2.0, frames, seed=3, random_msg=1, crc_len=8)
    assert 0 < int(crc_rows[1]) < int(crc_plain[1])
    again = dec.simulate(2.0, frames, seed=3, random_msg=1)
    assert np.array_equal(again, plain)


@pytest.mark.gpu
def test_device_monte_carlo_crc(code_files):
    """crc_len in nbl_simulate: CComm::CRCEncode on the device-drawn messages, CComm::CrcCheck on every decoded message.
    Pinned to the reference from both sides: code words the REFERENCE transmitted (golden fixture, CRC-8 inside) pass the
    device check, and so do the device's own messages; with errors, the identity
    CRC passes == error-free frames + undetected frames holds and undetected frames are a ~2^-8 fraction."""
    from nbldpc_b200 import capi
    z, meta = load_golden("cfg1_qspa_gf16")
    assert meta["crc"] == 8
    code = load_product_code(code_files, meta["code"])
    dec = capi.Decoder(code, capi.METHOD_QSPA, 10, precision=capi.PREC_F32, max_batch=20000)
    pool = z["tx_code_sym"].astype(np.int32)
    c = dec.simulate(9.0, 4800, seed=5, codewords=pool, crc_len=8)
    assert int(c[1]) == 0 and int(c[7]) == 4800, c
    c = dec.simulate(9.0, 4800, seed=5, random_msg=1, crc_len=8)
    assert int(c[1]) == 0 and int(c[7]) == 4800, c
    c = dec.simulate(9.0, 4800, seed=5, random_msg=1, crc_len=0)
    assert int(c[7]) == 0
    c = dec.simulate(1.0, 20000, seed=5, random_msg=1, crc_len=8)
    frames, err, und, ok = int(c[0]), int(c[1]), int(c[5]), int(c[7])
    assert err > 5000 and ok == frames - err + und, c
    assert und < err / 40, c                                  # about err / 256
    for crc in (16, 24):
        c = dec.simulate(9.0, 2000, seed=5, random_msg=1, crc_len=crc)
        assert int(c[1]) == 0
        assert int(c[7]) == (2000 if crc == 16 else 0)        # crcLen 24: CRC24A on encode, CRC24B on check, as the reference


@pytest.mark.gpu
def test_gpu_osd_order2_on_the_headline_code(code_files):
    """(512,256) GF(256), OSD order 2 with CRC-16 rows: 32 897 candidates per frame unranked from a linear index on the GPU
    against the oracle's nested loops (a few seconds per frame on the CPU)"""
    from nbldpc_b200 import capi
    from oracle import pyoracle as po
    z, meta = load_golden("f3_osd1_tems_gf256")
    code = load_product_code(code_files, meta["code"])
    lch = golden_lch(z, meta)[:3]
    oc = po.OracleCode(get_code_dict(meta["code"]))
    osd = po.OracleOsd(oc, 16, 9)
    want = [oc.decode(4, 3, 2, 2, 1.0, 0.5, l, osd, 2, 0.8, 0) for l in lch]
    assert all(w[0] == 0 for w in want)
    dec = capi.Decoder(code, 4, 3, 2, 2, 1.0, 0.5, precision=capi.PREC_F64, max_batch=len(lch))
    dec.enable_osd(2, 0.8, 0, 16, 9)
    r = dec.decode(lch)
    assert np.array_equal(r["sym"], np.stack([w[2] for w in want]))
    plain = capi.Decoder(code, 4, 3, 2, 2, 1.0, 0.5, precision=capi.PREC_F64, max_batch=len(lch)).decode(lch)
    assert (plain["sym"] != r["sym"]).any()
