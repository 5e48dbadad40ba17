"""the fp32 fast path and the fp64 parity path on the SAME frames (same seed => same messages and noise): error counters side
by side (development aid; headline configuration with the reference's early stop)"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from nbldpc_b200 import capi
from config_sweep import write_code_files
os.makedirs("/tmp/nbldpc_sweep", exist_ok=True)
paths, cons = write_code_files("/tmp/nbldpc_sweep")
code = capi.Code.from_file(paths["divsalar.UNBLDPC.512.256.GF.256"])
B, frames = 9472, 9472 * 40
out = {}
for name, prec, tie in (("fp32 fast (tems_warp)", capi.PREC_F32, 0), ("fp64 fast order (tems_warp)", capi.PREC_F64, 1), ("fp64 reference order (group)", capi.PREC_F64, 0)):
    dec = capi.Decoder(code, capi.METHOD_TEMS, 10, 2, 2, 1.0, 0.5, precision=prec, early_stop=1, max_batch=B, tie_mode=tie)
    c = dec.simulate(2.5, frames, seed=31, random_msg=1, crc_len=8)
    out[name] = [int(x) for x in c]
    print("%-30s %-10s frames %d  err frames %d  err syms %d  err bits %d  converged %d  undetected %d  iterations %d" % ((name, dec.kernel()) + tuple(int(x) for x in c[:7])), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fp32_vs_fp64.json"), "w"), indent=1)
