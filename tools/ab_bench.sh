#!/bin/bash
# development aid (run on the GPU box): bench every variants/lib_*.so given, print value / kernel ms / FER
for v in "$@"; do
  NBLDPC_LIB=variants/lib_$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bv_$v.json 2> gpurun_out/bv_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open("gpurun_out/bv_%s.json"%v)); print(v, "value %.2f e2e %.2f kernel_ms %.3f early %.2f fer %.6f"%(d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["early_stop"]["value"], d["fer"]))
except Exception as e: print(v, "failed", e)
PY
done
