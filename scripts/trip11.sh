mkdir -p gpurun_out
bash scripts/gpu_tests.sh t11 conv ops model
python scripts/trace_conv.py c1x1_64_64_320 > gpurun_out/trace_c1x1_64b.log 2>&1
python scripts/bench_layers.py --json gpurun_out/layers_r1e.json > gpurun_out/layers_r1e.log 2>&1; cat gpurun_out/layers_r1e.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/bench_profile_r1c.json > gpurun_out/bench_r1c.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r1c.log | cut -c1-300
