"""The C++ harness (nbldpc_b200/host: CSimulation / CComm / CNBLDPC with the reference's call surface) against the
golden vectors dumped from the real reference: same seed => same transmitted symbols, same received points, same
channel LLRs (CPU part, no GPU needed) and -- on the GPU -- the same decoder outputs through CNBLDPC::Decoding."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bundle, get_code_dict, load_golden
from oracle import pyoracle as po
from oracle.refprofile import profile_text

CASES = ["cfg1_qspa_gf16", "cfg1_qspa_gf16_allzero", "cfg2_ems_gf64", "cfg3_tems_gf256", "cfg4_qspa_gf256_qam", "cfg5_tems_hbit256", "f2_bstems_gf64_nm9"]


def _workdir(tmp_path, meta, extra=None):
    b = bundle()
    code_file = meta["code"] + ".txt"
    (tmp_path / code_file).write_text(po.code_text(get_code_dict(meta["code"])))
    for name, pts in b["constellations"].items():
        (tmp_path / (name + ".txt")).write_text("".join("Point:\t%d\tReal:\t%r\tImag:\t%r\n" % (i, p[0], p[1]) for i, p in enumerate(pts)))
    cfg = dict(GFq=meta["q"], code_file=code_file, method=meta["method"], snr=meta["snr_db"], seed=meta["seed"], random_msg=meta["random_msg"],
               nqam=meta["nqam"], constellation=meta["constellation"], crc=meta["crc"], max_iter=meta["max_iter"])
    if meta["method"] == 2:
        cfg.update(ems_nm=meta["p0"], ems_nc=meta["p1"], ems_factor=meta["factor"], ems_offset=meta["offset"])
    if meta["method"] == 4:
        cfg.update(tems_nr=meta["p0"], tems_nc=meta["p1"], tems_factor=meta["factor"], tems_offset=meta["offset"])
    if meta["method"] == 7:
        cfg.update(bs_nm=meta["p0"], bs_nc=meta["p1"], bs_factor=meta["factor"], bs_offset=meta["offset"])
    cfg.update(extra or {})
    (tmp_path / "prof.txt").write_text(profile_text(cfg))
    return cfg


def _run_dump(tmp_path, frames, decode):
    from nbldpc_b200 import build
    exe = [p for p in build.build_host() if p.endswith("harness_dump")][0]
    subprocess.check_call([exe, "prof.txt", str(frames), "out.bin", "1" if decode else "0"], cwd=tmp_path)
    raw = (tmp_path / "out.bin").read_bytes()
    _, F, N, q, max_iter, L, K, method = struct.unpack_from("8i", raw, 0)
    (sigma,) = struct.unpack_from("d", raw, 32)
    off, recs = 40, []
    for _ in range(F):
        r = {}
        r["rx"] = np.frombuffer(raw, np.float64, 2 * L, off).reshape(L, 2); off += 16 * L
        r["tx"] = np.frombuffer(raw, np.int32, N, off); off += 4 * N
        r["msg"] = np.frombuffer(raw, np.int32, K, off); off += 4 * K
        r["ret"], r["iters"] = struct.unpack_from("2i", raw, off); off += 8
        r["dec"] = np.frombuffer(raw, np.int32, N, off); off += 4 * N
        r["lch"] = np.frombuffer(raw, np.float64, N * (q - 1), off).reshape(N, q - 1); off += 8 * N * (q - 1)
        r["post"] = np.frombuffer(raw, np.float64, N * (q - 1), off).reshape(N, q - 1); off += 8 * N * (q - 1)
        recs.append(r)
    assert off == len(raw)
    return sigma, recs


@pytest.mark.parametrize("name", CASES)
def test_harness_feeds_the_decoder_the_reference_inputs(name, tmp_path):
    """PN source, CRC, encoder, bit packing, mapping, the reference RNG and the demapper: bit-identical to the reference"""
    z, meta = load_golden(name)
    _workdir(tmp_path, meta)
    F = min(len(z["ret"]), 6)
    sigma, recs = _run_dump(tmp_path, F, decode=False)
    assert sigma == meta["sigma"]
    for f, r in enumerate(recs):
        assert np.array_equal(r["tx"], z["tx_code_sym"][f]), "encoder output differs"
        assert np.array_equal(r["msg"], z["tx_msg_sym"][f])
        assert np.array_equal(r["rx"], z["rx"][f]), "received points differ (RNG / mapping)"
        assert np.array_equal(r["lch"].sum(axis=1), z["lch_sum"][f])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_harness_decoding_matches_reference(name, tmp_path):
    """CComm::Transmission() through CNBLDPC::Decoding on the GPU (fp64 parity mode) == reference outputs"""
    z, meta = load_golden(name)
    _workdir(tmp_path, meta)
    F = len(z["ret"])
    _, recs = _run_dump(tmp_path, F, decode=True)
    for f, r in enumerate(recs):
        assert (r["ret"], r["iters"]) == (int(z["ret"][f]), int(z["iters"][f])), f
        assert np.array_equal(r["dec"], z["dec_sym"][f])
        ref = z["lpost_sample"][f]
        sel = ref > -25.0 if meta["method"] == 1 else np.ones_like(ref, bool)
        err = np.abs(r["post"][:, z["lpost_idx"]] - ref)[sel] / np.maximum(np.abs(ref[sel]), 1.0)
        assert err.max() < 1e-6


@pytest.mark.gpu
def test_simulation_driver_reproduces_reference_fer_point(tmp_path):
    """nbldpc_sim (batched driver) on the survey's QSPA GF(16) anchor: 3 dB, all-zero codeword, seed 1, parallel 1 ->
    the reference reports 51 frame errors in 3380 frames.  Same RNG stream and stop rule => same row."""
    from nbldpc_b200 import build
    z, meta = load_golden("cfg1_qspa_gf16_allzero")
    _workdir(tmp_path, meta, dict(snr=3.0, min_err=50, min_cycle=2000, random_msg=0))
    exe = [p for p in build.build_host() if p.endswith("nbldpc_sim")][0]
    out = subprocess.run([exe, "prof.txt"], cwd=tmp_path, capture_output=True, text=True, timeout=600).stdout
    rows = [l for l in out.replace("\r", "\n").splitlines() if l.startswith("3\t")]
    last = rows[-1].split("\t")
    # the row the compiled reference prints for this profile (oracle/_ref/nbldpc_ref, 83 s on one CPU core):
    #   3  51  0  1.872226e-03  5.491864e-03  1.508876e-02  0.000000e+00
    assert last[:7] == ["3", "51", "0", "1.872226e-03", "5.491864e-03", "1.508876e-02", "0.000000e+00"], rows[-1]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["f3_osd2_tems_gf16", "f3_osd1_crc_ems_gf16", "f3_osd_alone_gf64"])
def test_harness_osd_matches_reference(name, tmp_path):
    """profile keys OSD_order / OSD_factor / OSD_flag / crc_correctLen and decode method 6 through CNBLDPC::Decoding:
    failed frames come back with the reference's OSD decision"""
    z, meta = load_golden(name)
    _workdir(tmp_path, meta, dict(osd_order=meta["osd_order"], osd_factor=meta["osd_factor"], osd_flag=meta["osd_flag"],
                                  crc_correct=meta["crc_correct"]))
    F = len(z["ret"])
    _, recs = _run_dump(tmp_path, F, decode=True)
    assert (z["ret"] == 0).any()
    for f, r in enumerate(recs):
        assert r["ret"] == int(z["ret"][f]), f
        if meta["method"] != 6:
            assert r["iters"] == int(z["iters"][f]), f
        assert np.array_equal(r["dec"], z["dec_sym"][f]), f
