mkdir -p gpurun_out
bash scripts/gpu_tests.sh t5 conv ops
python scripts/bench_layers.py --json gpurun_out/layers_r1b.json > gpurun_out/layers_r1b.log 2>&1; cat gpurun_out/layers_r1b.log
