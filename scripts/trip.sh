mkdir -p gpurun_out
bash scripts/gpu_tests.sh t39 ops model fullsize
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --profile-json gpurun_out/bench_profile_small2.json > gpurun_out/bench_small2.log 2> gpurun_out/bench_small2.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_small2.log') if x.startswith('{')]
d=json.loads(l[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
for k,v in d['kernels'].items(): print(k, v['ms_per_step'], v.get('gbs'))
pl=json.load(open('gpurun_out/bench_profile_small2.json'))['per_launch']
for n,ms,m in pl:
    if 'focus' in n or 'mp' in n: print(n, ms)
PY
