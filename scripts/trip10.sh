mkdir -p gpurun_out
bash scripts/gpu_tests.sh t10 conv
python scripts/trace_conv.py c3x3_focus_16_64_640 > gpurun_out/trace_focus.log 2>&1
python scripts/trace_conv.py c1x1_64_64_320 > gpurun_out/trace_c1x1_64.log 2>&1
python scripts/bench_layers.py --json gpurun_out/layers_r1d.json > gpurun_out/layers_r1d.log 2>&1; cat gpurun_out/layers_r1d.log
