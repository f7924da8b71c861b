mkdir -p gpurun_out
SKB_CONV_PAIR=0 python scripts/trace_conv.py c3x3_128_160 > gpurun_out/trace_3x3_p0.log 2>&1
SKB_CONV_PAIR=1 python scripts/trace_conv.py c3x3_128_160 > gpurun_out/trace_3x3_p1.log 2>&1
echo done
