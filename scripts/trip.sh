mkdir -p gpurun_out
bash scripts/gpu_tests.sh t34 conv ops model
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_ab_$tag.log 2> gpurun_out/bench_ab_$tag.err; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_ab_$tag.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('$tag', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], 'attn', d['kernels']['attention']['ms_per_step'], 'conv', d['kernels']['conv']['ms_per_step'])
else:
    print('$tag', 'no line'); print(open('gpurun_out/bench_ab_$tag.err').read()[-800:])
PY
}
run new1 A=1
run old1 SKB_CONV_PAIR=0 SKB_ATT_NQ=1
run nq2only SKB_CONV_PAIR=0
run paironly SKB_ATT_NQ=1
run new2 A=1
run old2 SKB_CONV_PAIR=0 SKB_ATT_NQ=1
