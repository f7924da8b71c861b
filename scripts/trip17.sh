mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --tb=short > gpurun_out/t17_conv.log 2>&1; echo "conv tests exit $?"; tail -15 gpurun_out/t17_conv.log
echo HALO0; SKB_CONV_HALO=0 timeout 300 python scripts/bench_layers.py --only c3x3_128_160,c3x3_256_80,c3x3_512_40,c3x3_64_320 2>&1
echo HALO1; timeout 300 python scripts/bench_layers.py --only c3x3_128_160,c3x3_256_80,c3x3_512_40,c3x3_64_320 2>&1
