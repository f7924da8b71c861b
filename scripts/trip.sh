mkdir -p gpurun_out
L=gpurun_out/attn_v4.log; : > $L
SKB_ATT_V2=1 timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "flash" 2>&1 | tail -3 >> $L
run() { echo "== $*" >> $L; env "$@" timeout 300 python scripts/bench_layers.py --only attn_p3,attn_p4,attn_p5 --iters 7 2>&1 | grep attn_ >> $L; }
run SKB_ATT_V2=0
run SKB_ATT_V2=1
run SKB_ATT_V2=1 SKB_ATT_POLY=0
run SKB_ATT_V2=1 SKB_ATT_POLY=16
run SKB_ATT_V2=0
cat $L
python scripts/bench_nms_stress.py > gpurun_out/nms_stress.json 2> gpurun_out/nms_stress.err; cat gpurun_out/nms_stress.json; tail -3 gpurun_out/nms_stress.err
