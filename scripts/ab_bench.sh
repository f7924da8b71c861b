#!/bin/bash
# In-step A/B of tuning knobs on ONE box: each argument is an env assignment list (quote it), e.g.
#   bash scripts/ab_bench.sh "SKB_ATT_NQ=1" "SKB_ATT_NQ=2 SKB_ATT_POLY=16"
# Prints images/s, attention TF/s and conv TF/s of `bench.py --steps 10` for each setting (box-to-box spread is +-3 %).
for cfg in "$@"; do
  out=$(env $cfg python bench.py --steps 10 --warmup 3 --no-tiled --no-cpu-baseline --no-latency 2>/dev/null | tail -1)
  python - "$cfg" "$out" <<'PY'
import json, sys
cfg, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    k = d["kernels"]
    print(f"{cfg:40s} {d['value']:7.1f} img/s  e2e {d['e2e']['value']:7.1f}  attn {k['attention']['tflops']:6.1f} TF/s {k['attention']['ms_per_step']:6.2f} ms  conv {k['conv']['tflops']:6.1f} TF/s {k['conv']['ms_per_step']:6.2f} ms  clk {d['clocks']['sm_mhz']}")
except Exception as e:
    print(cfg, "FAILED", e, line[:200])
PY
done
