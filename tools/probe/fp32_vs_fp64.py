"""the fp32 fast path and the fp64 parity path on the SAME frames (same seed => same messages and noise): error counters side
by side (development aid; headline configuration with the reference's early stop)"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from nbldpc_b200 import capi
from config_sweep import write_code_files
os.makedirs("/tmp/nbldpc_sweep", exist_ok=True)
paths, cons = write_code_files("/tmp/nbldpc_sweep")
CASES = [("cfg3 T-EMS GF(256) 2.5 dB", "divsalar.UNBLDPC.512.256.GF.256", capi.METHOD_TEMS, 2, 2, 0.5, 2.5, 9472, 40),
         ("cfg2 EMS GF(64) 2 dB", "BDS.576.288.GF.64", capi.METHOD_EMS, 8, 3, 0.5, 2.0, 9472, 20),
         ("cfg1 QSPA GF(16) 3 dB", "divsalar.UNBLDPC.128.64.GF.16", capi.METHOD_QSPA, 0, 0, 0.0, 3.0, 37888, 10),
         ("QSPA GF(256) BPSK 2.5 dB", "divsalar.CNBLDPC.512.256.GF.256", capi.METHOD_QSPA, 0, 0, 0.0, 2.5, 2368, 20)]
only = sys.argv[1:]
out = {}
for cname, cf, method, p0, p1, off, snr, B, rounds in CASES:
    if only and not any(o in cname for o in only):
        continue
    code = capi.Code.from_file(paths[cf])
    frames = B * rounds
    modes = [("fp32", capi.PREC_F32, 0), ("fp64", capi.PREC_F64, 1 if method == capi.METHOD_TEMS else 0)]
    if method == capi.METHOD_TEMS:
        modes.append(("fp64 reference order", capi.PREC_F64, 0))
    for name, prec, tie in modes:
        dec = capi.Decoder(code, method, 10, p0, p1, 1.0, off, precision=prec, early_stop=1, max_batch=B, tie_mode=tie)
        c = dec.simulate(snr, frames, seed=31, random_msg=1, crc_len=8)
        out[cname + " / " + name] = dict(kernel=dec.kernel(), counters=[int(x) for x in c])
        print("%-26s %-22s %-10s frames %d  err frames %d  err syms %d  err bits %d  converged %d  iterations %d" % ((cname, name, dec.kernel()) + tuple(int(x) for x in c[:5]) + (int(c[6]),)), flush=True)
        del dec
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fp32_vs_fp64.json"), "w"), indent=1)
