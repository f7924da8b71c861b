mkdir -p gpurun_out
bash scripts/gpu_tests.sh t7 conv ops
python scripts/bench_layers.py --json gpurun_out/layers_r1c.json > gpurun_out/layers_r1c.log 2>&1; cat gpurun_out/layers_r1c.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/bench_profile3.json > gpurun_out/bench3.log 2>&1; echo "bench exit $?"; tail -2 gpurun_out/bench3.log | cut -c1-300
