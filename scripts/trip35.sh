for P in 0 8; do SKB_ATT_POLY=$P timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_poly$P.log 2>&1; python - <<PY
import json
l=json.loads(open('gpurun_out/bench_poly$P.log').read().strip().splitlines()[-1]); print('POLY $P', l['value'], l['kernels']['attention'], l['clocks'])
PY
done
