mkdir -p gpurun_out
bash scripts/gpu_tests.sh t34 conv model
timeout 300 python scripts/bench_layers.py 2>&1 | grep -v attn
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/bench_profile_r1k.json > gpurun_out/bench_r1k.log 2>&1; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r1k.log').read().strip().splitlines()[-1]); print('bench', l['value'], l['e2e']['value'], l['latency_b1']['p50_ms'], l['kernels']['conv'])
PY
