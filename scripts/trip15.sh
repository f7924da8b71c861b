mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --tb=short > gpurun_out/t15_conv.log 2>&1; echo "conv tests exit $?"; tail -15 gpurun_out/t15_conv.log
echo PAIR0; SKB_CONV_PAIR=0 timeout 300 python scripts/bench_layers.py 2>&1 | grep -v attn
echo PAIR1; timeout 300 python scripts/bench_layers.py 2>&1 | grep -v attn
