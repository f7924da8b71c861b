mkdir -p gpurun_out
bash scripts/gpu_tests.sh t26 nms model
timeout 900 python bench.py --variant skyeye_s --batch 32 --size 640 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s640.log 2>&1; echo "bench_s exit $?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_s640.log').read().strip().splitlines()[-1]); print('skyeye_s', l['value'], l['e2e']['value'], l['latency_b1']['p50_ms'], {k:v['ms_per_step'] for k,v in l['kernels'].items()})
PY
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1j.log 2>&1; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r1j.log').read().strip().splitlines()[-1]); print('skyeye_l', l['value'], l['e2e']['value'], l['latency_b1']['p50_ms'], l['kernels']['nms'])
PY
