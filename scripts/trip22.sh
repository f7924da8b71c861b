mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "flash" --tb=short > gpurun_out/t22_attn.log 2>&1; echo "attn tests exit $?"; tail -5 gpurun_out/t22_attn.log
echo V2; SKB_ATT_V=2 timeout 300 python scripts/bench_layers.py --only attn_p3,attn_p4,attn_p5 2>&1 | tail -3
echo V3; timeout 300 python scripts/bench_layers.py --only attn_p3,attn_p4,attn_p5 2>&1 | tail -3
