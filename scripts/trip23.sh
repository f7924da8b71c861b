mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.log 2>&1; echo "nograph exit $?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_nograph.log').read().strip().splitlines()[-1]); print('nograph', l['value'], l['e2e']['value'], l['latency_b1'])
PY
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "graph exit $?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_graph.log').read().strip().splitlines()[-1]); print('graph', l['value'], l['e2e']['value'], l['latency_b1'])
PY
