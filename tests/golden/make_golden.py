"""Generate the golden fixtures tests/golden/*.npz from the REAL reference (run in the build container only).

For every case the unmodified reference (oracle/_ref, built by `make -C oracle ref`) is driven through its own
CComm stages and CNBLDPC::Decoding by oracle/ref_dump.cpp; this script then
  1. checks that the C oracle (oracle/nbldpc_oracle.c) reproduces the reference bit for bit on those frames
     (channel LLRs from the demapper, return code, iteration count, hard decisions and every L_post double), and
  2. stores the frames as small fixtures: received points, transmitted symbols, the reference's outputs and a
     fixed sample of L_post entries (the full arrays are 130 KB per GF(256) frame).
/root/reference is not needed to *use* the fixtures.
"""
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
BUNDLE = json.load(open(os.path.join(HERE, "codes.json")))

from oracle.refprofile import DEFAULTS, profile_text  # noqa: E402


# name -> config.  `code` is a key of codes.json (or "H_bit_GF256"); frames kept small: fixtures are committed.
CASES = {
    "cfg1_qspa_gf16": dict(code="divsalar.UNBLDPC.128.64.GF.16", GFq=16, method=1, snr=2.0, frames=48),
    "cfg1_qspa_gf16_allzero": dict(code="divsalar.UNBLDPC.128.64.GF.16", GFq=16, method=1, snr=3.0, frames=16, random_msg=0),
    "cfg2_ems_gf64": dict(code="BDS.576.288.GF.64", GFq=64, method=2, snr=1.5, frames=16),
    "cfg3_tems_gf256": dict(code="divsalar.UNBLDPC.512.256.GF.256", GFq=256, method=4, snr=2.5, frames=10),
    "cfg3_tems_gf256_lowsnr": dict(code="divsalar.UNBLDPC.512.256.GF.256", GFq=256, method=4, snr=1.25, frames=6),
    # 3 dB: deep in the waterfall of this code over 256-QAM, so that frames run 5 .. 10 iterations and two of the eight do not
    # converge (at 5 .. 13 dB every frame converges within 8); the reference's box-plus needs ~2.6 s per frame-iteration
    "cfg4_qspa_gf256_qam": dict(code="divsalar.CNBLDPC.512.256.GF.256", GFq=256, method=1, snr=3.0, frames=8,
                                nqam=256, constellation="GRAY_256QAM.txt", random_msg=0),
    "cfg5_tems_hbit256": dict(code="H_bit_GF256", GFq=256, method=4, snr=3.0, frames=24),
    "x_tems_gf16_dc5": dict(code="divsalar.UNBLDPC.128.64.GF.16", GFq=16, method=4, snr=2.0, frames=24, tems_nr=2, tems_nc=3),
    "x_tems_gf64": dict(code="BDS.576.288.GF.64", GFq=64, method=4, snr=1.5, frames=12, tems_nr=3, tems_nc=2, tems_offset=0.3),
    "x_ems_gf16_dc5": dict(code="divsalar.UNBLDPC.128.64.GF.16", GFq=16, method=2, snr=2.0, frames=24, ems_nm=6, ems_nc=2, ems_offset=0.2),
    "x_ems_gf256": dict(code="divsalar.UNBLDPC.128.64.GF.256", GFq=256, method=2, snr=2.5, frames=6, ems_nm=8, ems_nc=3),
    "x_qspa_gf64": dict(code="BDS.576.288.GF.64", GFq=64, method=1, snr=1.5, frames=3),
    "f2_bstems_gf256": dict(code="divsalar.UNBLDPC.512.256.GF.256", GFq=256, method=7, snr=2.75, frames=10, bs_nm=4, bs_nc=3, bs_offset=0.2),
    "f2_bstems_gf64_nm9": dict(code="BDS.576.288.GF.64", GFq=64, method=7, snr=2.0, frames=12, bs_nm=9, bs_nc=2, bs_factor=1.25, bs_offset=0.1),
    # 8(f)#3: OSD post-processing of failed frames (low SNR so that most frames fail), CRC-aided rows, method 6 stand-alone
    "f3_osd2_tems_gf16": dict(code="divsalar.UNBLDPC.128.64.GF.16", GFq=16, method=4, snr=1.0, frames=32, max_iter=5, osd_order=2, crc_correct=0),
    "f3_osd1_crc_ems_gf16": dict(code="divsalar.UNBLDPC.128.64.GF.16", GFq=16, method=2, snr=1.0, frames=24, max_iter=4, osd_order=1,
                                 osd_factor=0.5, crc=8, crc_correct=8, ems_nm=6, ems_nc=2),
    "f3_osd3_qspa_hbit256": dict(code="H_bit_GF256", GFq=256, method=1, snr=1.5, frames=8, max_iter=3, osd_order=3, crc=16, crc_correct=4),
    "f3_osd1_tems_gf256": dict(code="divsalar.UNBLDPC.512.256.GF.256", GFq=256, method=4, snr=1.5, frames=4, max_iter=4, osd_order=1,
                               crc=24, crc_correct=24, osd_flag=1),
    "f3_osd_alone_gf64": dict(code="BDS.576.288.GF.64", GFq=64, method=6, snr=3.0, frames=6, osd_order=1, crc=8, crc_correct=8),
    "f2_bstems_gf16": dict(code="divsalar.UNBLDPC.128.64.GF.16", GFq=16, method=7, snr=2.5, frames=24, bs_nm=3, bs_nc=3, bs_offset=0.0),
}


def get_code(name):
    if name in BUNDLE["codes"]:
        c = BUNDLE["codes"][name]
        return dict(N=c["N"], M=c["M"], q=c["q"], vdeg=[len(l) for l in c["var"]], cdeg=[len(l) for l in c["chk"]],
                    var=[[tuple(e) for e in l] for l in c["var"]], chk=[[tuple(e) for e in l] for l in c["chk"]])
    return po.hbit_to_code(BUNDLE["h_bit"][name], 256)


def run_case(name, cfg):
    code = get_code(cfg["code"])
    frames = cfg["frames"]
    with tempfile.TemporaryDirectory() as tmp:
        for f in os.listdir(os.path.join(REFDIR, "data")):
            os.symlink(os.path.join(REFDIR, "data", f), os.path.join(tmp, f))
        code_file = cfg["code"] + ".txt"
        if not os.path.exists(os.path.join(tmp, code_file)):      # H_bit image -> reference code format
            open(os.path.join(tmp, code_file), "w").write(po.code_text(code))
        pc = dict(cfg); pc["code_file"] = code_file
        open(os.path.join(tmp, "prof.txt"), "w").write(profile_text(pc))
        subprocess.check_call([os.path.join(REFDIR, "ref_dump"), "prof.txt", str(frames), "out.bin"], cwd=tmp,
                              stderr=subprocess.DEVNULL)
        raw = open(os.path.join(tmp, "out.bin"), "rb").read()
    magic, F, N, q, max_iter, L, K, method = struct.unpack_from("8i", raw, 0)
    assert magic == 0x4E424C44 and F == frames
    (sigma,) = struct.unpack_from("d", raw, 32)
    off = 40
    q1 = q - 1
    rx = np.zeros((F, L, 2)); tx = np.zeros((F, N), np.int32); msg = np.zeros((F, K), np.int32)
    ret = np.zeros(F, np.int32); iters = np.zeros(F, np.int32); dec = np.zeros((F, N), np.int32)
    lch = np.zeros((F, N, q1)); lpost = np.zeros((F, N, q1))

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype, count, off); off += a.nbytes
        return a
    for f in range(F):
        rx[f] = take(np.float64, 2 * L).reshape(L, 2)
        tx[f] = take(np.int32, N); msg[f] = take(np.int32, K)
        ret[f], iters[f] = take(np.int32, 2)
        dec[f] = take(np.int32, N)
        lch[f] = take(np.float64, N * q1).reshape(N, q1)
        lpost[f] = take(np.float64, N * q1).reshape(N, q1)
    assert off == len(raw)

    # --- pin the oracle against the reference, bit for bit ---
    d = dict(DEFAULTS); d.update(cfg)
    oc = po.OracleCode(code)
    osd = po.OracleOsd(oc, d["crc"], d["crc_correct"]) if d["osd_order"] >= 0 else None
    if method in (1, 6): p0, p1, fac, ofs = 0, 0, 1.0, 0.0
    elif method == 2: p0, p1, fac, ofs = d["ems_nm"], d["ems_nc"], d["ems_factor"], d["ems_offset"]
    elif method == 7: p0, p1, fac, ofs = d["bs_nm"], d["bs_nc"], d["bs_factor"], d["bs_offset"]
    else: p0, p1, fac, ofs = d["tems_nr"], d["tems_nc"], d["tems_factor"], d["tems_offset"]
    cons = None
    if d["nqam"] != 2:
        cons = np.array(BUNDLE["constellations"][d["constellation"][:-4]])
    for f in range(F):
        mine = po.demap_bpsk(q, N, rx[f, :, 0], sigma) if d["nqam"] == 2 else po.demap_qam(q, N, rx[f], cons, sigma)
        assert np.array_equal(mine, lch[f]), "oracle demapper differs from the reference"
        if method == 6:                                           # stand-alone OSD (Decoding_OSD_bit with flag 1): returns 0
            assert ret[f] == 0
            assert np.array_equal(osd.run(d["osd_order"], lch[f]), dec[f]), (name, f)
            continue
        r, it, out, post = oc.decode(method, max_iter, p0, p1, fac, ofs, lch[f], osd, d["osd_order"], d["osd_factor"], d["osd_flag"])
        assert (r, it) == (ret[f], iters[f]), (name, f, r, it, ret[f], iters[f])
        assert np.array_equal(out, dec[f]), (name, f, np.flatnonzero(out != dec[f]))
        assert np.array_equal(post, lpost[f]), (name, f, np.abs(post - lpost[f]).max())
    rng = np.random.RandomState(12345)
    idx = np.sort(rng.choice(q1, min(q1, 32), replace=False)).astype(np.int32)
    meta = dict(name=name, code=cfg["code"], q=q, N=N, K=K, method=method, max_iter=max_iter, p0=p0, p1=p1,
                factor=fac, offset=ofs, snr_db=d["snr"], sigma=sigma, seed=d["seed"], random_msg=d["random_msg"],
                nqam=d["nqam"], constellation=d["constellation"], crc=d["crc"], crc_correct=d["crc_correct"],
                osd_order=d["osd_order"], osd_factor=d["osd_factor"], osd_flag=d["osd_flag"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), rx=rx, tx_code_sym=tx.astype(np.int16),
                        tx_msg_sym=msg.astype(np.int16), ret=ret.astype(np.int8), iters=iters.astype(np.int8),
                        dec_sym=dec.astype(np.int16), lpost_idx=idx, lpost_sample=lpost[:, :, idx],
                        lch_sum=lch.sum(axis=2))
    print("%-28s frames %3d  converged %3d  iters %s  oracle == reference bit-exact" %
          (name, F, int(ret.sum()), np.bincount(iters, minlength=max_iter + 1)[1:].tolist()))


if __name__ == "__main__":
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so", "ref"])
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n, CASES[n])
