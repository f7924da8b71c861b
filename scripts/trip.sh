mkdir -p gpurun_out
bash scripts/gpu_tests.sh t38 ops nms model fullsize
python scripts/bench_nms_stress.py 2>/dev/null | cut -c1-400
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_small.log 2> gpurun_out/bench_small.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_small.log') if x.startswith('{')]
d=json.loads(l[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
for k,v in d['kernels'].items(): print(k, v['ms_per_step'], v.get('gbs'))
PY
