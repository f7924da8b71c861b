mkdir -p gpurun_out
bash scripts/gpu_tests.sh t28 conv model
for P in 0 1; do SKB_PDL=$P timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$P.log 2>&1; python - <<PY
import json
l=json.loads(open('gpurun_out/bench_pdl$P.log').read().strip().splitlines()[-1]); print('PDL $P', l['value'], l['e2e']['value'], l['latency_b1']['p50_ms'], l['kernels']['conv']['ms_per_step'])
PY
done
