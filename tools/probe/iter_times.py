#!/usr/bin/env python
"""Development probe: duration of each iteration kernel of one fixed-10 batch of the headline workload (ms).
    NBLDPC_LIB=variants/lib_x.so python tools/probe/iter_times.py [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from nbldpc_b200 import capi
from bench import bundle_code_text, CODE, TEMS
B = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
open("/tmp/it_code.txt", "w").write(bundle_code_text(CODE))
code = capi.Code.from_file("/tmp/it_code.txt")
for es in (0, 1):
    dec = capi.Decoder(code, precision=capi.PREC_F32, early_stop=es, max_batch=B, **TEMS)
    for rep in range(3):
        c = dec.simulate(2.5, B, seed=1 + rep, random_msg=1, crc_len=8)
    t = dec.iter_timing()
    act = dec.iter_active()
    print(os.environ.get("NBLDPC_LIB", "lib"), "early_stop", es, "iter ms:", " ".join("%.2f" % x for x in t), " sum %.2f" % sum(t), " FER %.4f" % (c[1] / c[0]))
    print("   active frames:", " ".join(str(x) for x in act))
    full = max(t[1:4])
    print("   ms per iteration if cost were proportional to the active frames:", " ".join("%.2f" % (full * a / B) for a in act[:len(t)]),
          " sum %.2f" % sum(full * a / B for a in act[:len(t)]))
