#!/usr/bin/env python
"""Secondary measurements (not the bench line): the other BASELINE.json configurations through the device Monte-Carlo path.

    python tools/config_sweep.py [--out profiles/r01_configs.json] [--cfg5-frames 100000000]

cfg1  divsalar.UNBLDPC.128.64.GF.16   QSPA           BPSK   Eb/N0 1..4 dB         FER/BER per point + throughput
cfg2  BDS.576.288.GF.64               EMS 8/3/0.5    BPSK   1.5..3.0 dB, early stop
cfg3  (bench.py)                      frames-in-flight sweep of the headline kernel, fixed 10 iterations
cfg4  divsalar.CNBLDPC.512.256.GF.256 QSPA (FWHT)    Gray 256-QAM 12..16 dB (symbol value -> constellation index)
cfg5  H_bit_GF256                     T-EMS 2/2/0.5  BPSK   low-FER Monte-Carlo with the reference stop rule

Every number is device time of whole nbl_simulate calls (CUDA events through torch on the decoder's stream), random
code words from the device encoder, fp32 fast mode.  Rooflines per SURVEY 8(d): B_iter = 4 q (4E + N) bytes per
frame-iteration against the measured HBM copy bandwidth; for early-stop runs with the measured mean iteration count.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_code_files(tmp):
    c = json.load(open(os.path.join(ROOT, "tests", "golden", "codes.json")))
    paths = {}
    for name, cd in c["codes"].items():
        out = ["%d %d %d" % (cd["N"], cd["M"], cd["q"]), "%d %d" % (max(map(len, cd["var"])), max(map(len, cd["chk"])))]
        out.append(" ".join(str(len(l)) for l in cd["var"]) + " ")
        out.append(" ".join(str(len(l)) for l in cd["chk"]) + " ")
        for l in cd["var"]:
            out.append(" ".join("%d %d" % (r + 1, g) for r, g in l) + " ")
        for l in cd["chk"]:
            out.append(" ".join("%d %d" % (v + 1, g) for v, g in l) + " ")
        p = os.path.join(tmp, name + ".txt")
        open(p, "w").write("\n".join(out) + "\n")
        paths[name] = p
    p = os.path.join(tmp, "H_bit_GF256.txt")
    open(p, "w").write("".join("\t".join(r) + "\t\n" for r in c["h_bit"]["H_bit_GF256"]))
    paths["H_bit_GF256"] = p
    return paths, c["constellations"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_configs.json"))
    ap.add_argument("--cfg5-frames", type=int, default=20_000_000)
    ap.add_argument("--skip", default="")
    a = ap.parse_args()
    import numpy as np
    import torch
    from nbldpc_b200 import capi, montecarlo as mc
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    tmp = "/tmp/nbldpc_sweep"
    os.makedirs(tmp, exist_ok=True)
    paths, cons = write_code_files(tmp)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    res = {"hbm_peak_gbs": peaks["hbm_gbs"], "note": "device time of nbl_simulate calls, fp32, random code words (device encoder)", "points": []}

    def run(tag, code, method, p0, p1, offset, ebn0, batch, steps, early_stop, modulation=2, constellation=None, max_iter=10):
        dec = capi.Decoder(code, method, max_iter, p0, p1, 1.0, offset, precision=capi.PREC_F32, early_stop=early_stop,
                           max_batch=batch, stream=stream.cuda_stream)
        cnt = torch.zeros(capi.NUM_COUNTERS, dtype=torch.int64, device="cuda")
        kw = dict(modulation=modulation, constellation=constellation, random_msg=1)
        for i in range(2):
            dec.simulate_async(ebn0, batch, cnt, seed=7, first_frame=i * batch, **kw)
        torch.cuda.synchronize()
        cnt.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            dec.simulate_async(ebn0, batch, cnt, seed=7, first_frame=(2 + i) * batch, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        c = cnt.cpu().numpy()
        frames = int(c[0])
        K = code.N - code.M
        info = frames * K * code.p / (ms * 1e-3) / 1e6
        mean_it = c[6] / max(1, frames)
        its = max_iter if not early_stop else mean_it
        b_iter = 4 * code.q * (4 * code.E + code.N)
        gbs = frames * its * b_iter / (ms * 1e-3) / 1e9
        pt = {"cfg": tag, "ebn0_db": ebn0, "frames": frames, "batch": batch, "early_stop": bool(early_stop), "info_mbit_s": info,
              "fer": c[1] / max(1, frames), "ber": c[3] / max(1, frames * K * code.p), "mean_iterations": float(mean_it),
              "algorithmic_gb_s": gbs, "hbm_frac": gbs / peaks["hbm_gbs"], "ms": ms}
        res["points"].append(pt)
        print(json.dumps(pt), flush=True)
        return pt

    skip = set(a.skip.split(","))
    if "cfg1" not in skip:
        code = capi.Code.from_file(paths["divsalar.UNBLDPC.128.64.GF.16"])
        for snr in (1.0, 2.0, 3.0, 4.0):
            run("cfg1 QSPA GF(16) (128,64) BPSK", code, capi.METHOD_QSPA, 0, 0, 0.0, snr, 148 * 512, 4, 1)
    if "cfg2" not in skip:
        code = capi.Code.from_file(paths["BDS.576.288.GF.64"])
        for snr in (1.5, 2.0, 2.5, 3.0):
            run("cfg2 EMS GF(64) (576,288) n_m=8 n_c=3", code, capi.METHOD_EMS, 8, 3, 0.5, snr, 148 * 128, 4, 1)
        run("cfg2 EMS GF(64), fixed 10 iterations", code, capi.METHOD_EMS, 8, 3, 0.5, 2.0, 148 * 128, 4, 0)
    if "cfg3" not in skip:
        code = capi.Code.from_file(paths["divsalar.UNBLDPC.512.256.GF.256"])
        for b in (148, 296, 592, 1184, 4736, 18944, 75776):
            run("cfg3 T-EMS GF(256) (512,256) frames in flight %d" % b, code, capi.METHOD_TEMS, 2, 2, 0.5, 2.5, b, max(1, min(20, 18944 // b)), 0)
    if "cfg4" not in skip:
        code = capi.Code.from_file(paths["divsalar.CNBLDPC.512.256.GF.256"])
        cs = np.array(cons["GRAY_256QAM"], np.float64)
        for snr in (12.0, 14.0, 16.0):
            run("cfg4 QSPA-FWHT GF(256) (512,256) 256-QAM", code, capi.METHOD_QSPA, 0, 0, 0.0, snr, 148 * 16, 3, 1, modulation=256, constellation=cs)
    if "cfg5" not in skip:
        code = capi.Code.from_hbit(paths["H_bit_GF256"], 256)
        batch = 148 * 512
        dec = capi.Decoder(code, capi.METHOD_TEMS, 10, 2, 2, 1.0, 0.5, precision=capi.PREC_F32, early_stop=1, max_batch=batch,
                           stream=stream.cuda_stream)
        dev = torch.zeros(capi.NUM_COUNTERS, dtype=torch.int64, device="cuda")
        for snr in (3.0, 4.0, 5.0, 6.0):
            def simulate(lo, hi):
                dev.zero_()
                dec.simulate_async(snr, hi - lo, dev, seed=11, first_frame=lo, random_msg=1)
                return dev.cpu().numpy()
            t0 = time.time()
            total, rounds = mc.run_point(simulate, lambda v: None, 0, 1, batch, mc.StopRule(min_frames=300, min_err_frames=100, max_frames=a.cfg5_frames))
            dt = time.time() - t0
            r = mc.rates(total, code.N - code.M, code.p)
            pt = {"cfg": "cfg5 H_bit_GF256 T-EMS Monte-Carlo, reference stop rule (100 frame errors or the frame cap)", "ebn0_db": snr,
                  "rounds": rounds, "wall_s": dt, "frames_per_s": r["frames"] / dt, **r}
            res["points"].append(pt)
            print(json.dumps(pt), flush=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
