"""CPU tests of the oracle (oracle/nbldpc_oracle.c) against the golden vectors dumped from the real reference,
and -- when the compiled reference is present (build container) -- against the reference run live."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, REFERENCE, bundle, get_code_dict, golden_lch, golden_names, load_golden
from oracle import pyoracle as po

FAST = [n for n in golden_names() if n not in ("cfg4_qspa_gf256_qam", "x_qspa_gf64")]   # QSPA at q>=64 is slow on a CPU


@pytest.mark.parametrize("name", FAST)
def test_oracle_reproduces_reference_golden(name):
    z, meta = load_golden(name)
    oc = po.OracleCode(get_code_dict(meta["code"]))
    F = min(len(z["ret"]), 8)
    lch = golden_lch(z, meta, F)
    assert np.allclose(lch.sum(axis=2), z["lch_sum"][:F], rtol=0, atol=1e-9)
    for f in range(F):
        ret, it, out, post = oc.decode(meta["method"], meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"], lch[f])
        assert (ret, it) == (int(z["ret"][f]), int(z["iters"][f]))
        assert np.array_equal(out, z["dec_sym"][f])
        assert np.array_equal(post[:, z["lpost_idx"]], z["lpost_sample"][f])     # bit-exact, not approximately


def test_oracle_qspa_gf256_qam_one_frame():
    """the frame of the 256-QAM fixture that needs the fewest iterations (the oracle's box-plus, like the reference's, takes ~2.6 s
    per frame-iteration); tests/golden/make_golden.py pinned all eight, including the two that do not converge"""
    z, meta = load_golden("cfg4_qspa_gf256_qam")
    oc = po.OracleCode(get_code_dict(meta["code"]))
    f = int(np.argmin(z["iters"]))
    lch = golden_lch(z, meta)[f]
    ret, it, out, post = oc.decode(1, meta["max_iter"], 0, 0, 1.0, 0.0, lch)
    assert (ret, it) == (int(z["ret"][f]), int(z["iters"][f])) and np.array_equal(out, z["dec_sym"][f])
    assert np.array_equal(post[:, z["lpost_idx"]], z["lpost_sample"][f])


def test_gf_tables_properties():
    """SURVEY section 4 fixture 1: add == xor, mul == carry-less product mod the polynomial, inverses"""
    for q, poly in po.PRIM_POLY.items():
        if q > 256:
            continue
        mul, inv = po.gf_tables(q)
        assert (mul[1] == np.arange(q)).all() and (mul[:, 0] == 0).all() and (mul == mul.T).all()
        for x in range(1, q):
            assert mul[x, inv[x]] == 1
            assert sorted(mul[x]) == list(range(q))
        assert inv[0] == 0
        a, b, c = 3 % q, 5 % q, (q - 1)
        assert mul[a, b ^ c] == mul[a, b] ^ mul[a, c]


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
def test_gf_tables_equal_reference_files():
    for q in (4, 8, 16, 32, 64, 128, 256):
        tok = open(os.path.join(REFERENCE, "Arith.Table.GF.%d.txt" % q)).read().replace(".", " ").split()
        i = tok.index("Table:") + 1
        mul_ref = np.array(tok[i:i + q * q], int).reshape(q, q)
        mul, inv = po.gf_tables(q)
        assert np.array_equal(mul, mul_ref)
        j = len(tok) - q
        assert np.array_equal(inv, np.array(tok[j:], int))


def test_code_bundle_invariants():
    """SURVEY section 4 fixture 3 on every bundled code: both edge lists agree, links ascending, no multi-edges"""
    for name in bundle()["codes"]:
        c = get_code_dict(name)
        edges_v = {(v, r): g for v, l in enumerate(c["var"]) for r, g in l}
        edges_c = {(v, r): g for r, l in enumerate(c["chk"]) for v, g in l}
        assert edges_v == edges_c, name
        assert sum(c["vdeg"]) == sum(c["cdeg"]) == len(edges_v)
        for l in c["var"] + c["chk"]:
            idx = [i for i, _ in l]
            assert idx == sorted(set(idx))
        assert po.parse_code_text(po.code_text(c))["chk"] == c["chk"]


def test_hbit_gf256_image_decodes_to_survey_edge_list():
    c = get_code_dict("H_bit_GF256")
    assert (c["N"], c["M"]) == (16, 8)
    assert c["var"][0] == [(0, 98), (6, 229)] and c["var"][15] == [(0, 1), (7, 1)] and c["var"][9] == [(1, 58), (5, 204)]


def test_constellations_unit_energy_and_gray():
    """SURVEY section 4 fixture 6"""
    for name, m in (("GRAY_64QAM", 64), ("GRAY_256QAM", 256)):
        pts = np.array(bundle()["constellations"][name])
        assert pts.shape == (m, 2)
        assert abs((pts ** 2).sum(axis=1).mean() - 1.0) < 1e-3
        d = np.linalg.norm(pts[:, None] - pts[None], axis=2) + np.eye(m) * 9
        dmin = d.min()
        for i in range(m):
            for j in np.where(d[i] < dmin * 1.01)[0]:
                assert bin(i ^ int(j)).count("1") == 1


def test_noiseless_all_zero_converges_at_iteration_one():
    for code_name, method, p0, p1 in [("divsalar.UNBLDPC.128.64.GF.16", 1, 0, 0), ("BDS.576.288.GF.64", 2, 8, 3),
                                      ("divsalar.UNBLDPC.128.64.GF.256", 4, 2, 2)]:
        c = get_code_dict(code_name)
        oc = po.OracleCode(c)
        ret, it, out, _ = oc.decode(method, 10, p0, p1, 1.0, 0.5, np.full((c["N"], c["q"] - 1), -10.0))
        assert (ret, it) == (1, 1) and not out.any()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_dump")), reason="compiled reference not built")
def test_oracle_against_live_reference(tmp_path):
    """fresh frames (not in the fixtures) from the compiled reference vs the oracle, bit for bit"""
    import struct
    from oracle import refprofile as mg
    ref = os.path.join(ROOT, "oracle", "_ref")
    for f in os.listdir(os.path.join(ref, "data")):
        os.symlink(os.path.join(ref, "data", f), tmp_path / f)
    cfg = dict(code_file="divsalar.UNBLDPC.128.64.GF.16.txt", GFq=16, method=4, snr=1.5, seed=7, tems_nr=2, tems_nc=2)
    (tmp_path / "prof.txt").write_text(mg.profile_text(cfg))
    subprocess.check_call([os.path.join(ref, "ref_dump"), "prof.txt", "6", "out.bin", "3"], cwd=tmp_path, stderr=subprocess.DEVNULL)
    raw = (tmp_path / "out.bin").read_bytes()
    _, F, N, q, max_iter, L, K, method = struct.unpack_from("8i", raw, 0)
    off = 40
    oc = po.OracleCode(get_code_dict("divsalar.UNBLDPC.128.64.GF.16"))
    for _ in range(F):
        off += 16 * L + 4 * N + 4 * K
        ret, iters = struct.unpack_from("2i", raw, off); off += 8
        dec = np.frombuffer(raw, np.int32, N, off); off += 4 * N
        lch = np.frombuffer(raw, np.float64, N * (q - 1), off).reshape(N, q - 1); off += 8 * N * (q - 1)
        lpost = np.frombuffer(raw, np.float64, N * (q - 1), off).reshape(N, q - 1); off += 8 * N * (q - 1)
        r, it, out, post = oc.decode(4, max_iter, 2, 2, 1.0, 0.5, lch)
        assert (r, it) == (ret, iters) and np.array_equal(out, dec) and np.array_equal(post, lpost)


def test_ems_fresh_sum_switch_is_an_analysis_aid_only():
    """orc_set_ems_fresh_sums (off by default) forms every EMS configuration sum afresh instead of in the reference's running
    add / subtract sum: on the golden EMS frames it changes posteriors by rounding residue only (same decisions, same iteration
    counts), and switching it off again restores the reference's arithmetic bit for bit."""
    name = "cfg2_ems_gf64"
    z, meta = load_golden(name)
    oc = po.OracleCode(get_code_dict(meta["code"]))
    lch = golden_lch(z, meta, 4)
    base = [oc.decode(meta["method"], meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"], l) for l in lch]
    po.set_ems_fresh_sums(1)
    try:
        fresh = [oc.decode(meta["method"], meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"], l) for l in lch]
    finally:
        po.set_ems_fresh_sums(0)
    again = [oc.decode(meta["method"], meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"], l) for l in lch]
    for b, f, a in zip(base, fresh, again):
        assert (b[0], b[1]) == (f[0], f[1]) and np.array_equal(b[2], f[2])
        assert np.abs(b[3] - f[3]).max() < 1e-9
        assert np.array_equal(b[3], a[3])
