#!/usr/bin/env python
"""Development probe: throughput of the headline configuration with the reference's tie order (tie_mode REFERENCE) on the
one-warp-per-row kernel against the group kernel, fp32 and fp64; error counters of every run on the same frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import time
from nbldpc_b200 import capi
from bench import bundle_code_text, CODE, TEMS
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
open("/tmp/rt_code.txt", "w").write(bundle_code_text(CODE))
code = capi.Code.from_file("/tmp/rt_code.txt")
for label, prec, tie, kern in (("fp32 fast order, warp kernel", capi.PREC_F32, capi.TIES_FAST, capi.KERNEL_AUTO),
                               ("fp32 reference order, warp kernel", capi.PREC_F32, capi.TIES_REFERENCE, capi.KERNEL_AUTO),
                               ("fp32 reference order, group kernel", capi.PREC_F32, capi.TIES_REFERENCE, capi.KERNEL_GROUP),
                               ("fp64 default (reference order), warp kernel", capi.PREC_F64, capi.TIES_AUTO, capi.KERNEL_AUTO),
                               ("fp64 fast order, warp kernel", capi.PREC_F64, capi.TIES_FAST, capi.KERNEL_AUTO),
                               ("fp64 reference order, group kernel", capi.PREC_F64, capi.TIES_AUTO, capi.KERNEL_GROUP)):
    dec = capi.Decoder(code, precision=prec, early_stop=1, max_batch=B, tie_mode=tie, kernel=kern, **TEMS)
    dec.simulate(2.5, B, seed=1, random_msg=1, crc_len=8)
    t0 = time.perf_counter()
    c = dec.simulate(2.5, B, seed=2, random_msg=1, crc_len=8)
    dt = time.perf_counter() - t0
    print("%-46s %-10s %8.2f Mbit/s  counters %s" % (label, dec.kernel(), B * 256 / dt / 1e6, c.tolist()))
