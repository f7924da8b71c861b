mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -m gpu -q -k "flash or attention" 2>&1 | tail -2
L=gpurun_out/attn_x4.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 300 python scripts/bench_layers.py --only attn_p3,attn_p4,attn_p5 --iters 7 2>&1 | grep attn_ >> $L; }
run SKB_ATT_X4=0
run SKB_ATT_X4=1
run SKB_ATT_X4=0 SKB_ATT_NQ=1
run SKB_ATT_X4=1 SKB_ATT_NQ=1
run SKB_ATT_X4=1 SKB_ATT_POLY=16
run SKB_ATT_X4=1 SKB_ATT_POLY=0
cat $L
SKB_ATT_PROF=1 python scripts/attn_prof.py
runb() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_ab_$tag.log 2> gpurun_out/bench_ab_$tag.err; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_ab_$tag.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('$tag', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], 'attn', d['kernels']['attention']['ms_per_step'], d['roofline']['frac'])
else:
    print('$tag', 'no line'); print(open('gpurun_out/bench_ab_$tag.err').read()[-800:])
PY
}
runb x4_0 SKB_ATT_X4=0
runb x4_1 SKB_ATT_X4=1
runb x4_0b SKB_ATT_X4=0
runb x4_1b SKB_ATT_X4=1
