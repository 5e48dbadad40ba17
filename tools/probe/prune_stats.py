#!/usr/bin/env python
"""Development probe (needs a library built with -DTW_STATS=1: tools/build_variant.sh stats -DNBL_DEV_MINIMAL -DTW_STATS=1):
per iteration of the headline configuration, how many of the 255 pair-search entries per check row the value bound keeps.
    NBLDPC_LIB=variants/lib_stats.so python tools/probe/prune_stats.py"""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nbldpc_b200 import capi
from bench import bundle_code_text, CODE, TEMS

open("/tmp/prune_code.txt", "w").write(bundle_code_text(CODE))
code = capi.Code.from_file("/tmp/prune_code.txt")
L = capi.lib()
buf = (ctypes.c_ulonglong * 128)()
out = {}
for es in (0, 1):
    dec = capi.Decoder(code, precision=capi.PREC_F32, early_stop=es, max_batch=2368, **TEMS)
    L.nbl_dev_stats(buf, 1)
    c = dec.simulate(2.5, 2368, seed=1, random_msg=1, crc_len=8)
    L.nbl_dev_stats(buf, 1)
    st = np.array(buf[:], dtype=np.float64).reshape(16, 8)
    rows = []
    for it in range(1, 10):
        n, live, second, blocks = st[it][:4]
        if n:
            rows.append(dict(iteration=it, check_rows=int(n), live_entries_per_row=live / n, second_kind_per_row=second / n, live_blocks_per_row=blocks / n))
            print("early_stop %d iter %d rows %7d  live entries %.1f of 255  second kind %.2f  blocks with a live entry %.1f of 32" % (es, it, n, live / n, second / n, blocks / n))
    out["early_stop_%d" % es] = dict(fer=float(c[1]) / float(c[0]), per_iteration=rows)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "prune_stats.json"), "w"), indent=1)
