mkdir -p gpurun_out
bash scripts/gpu_tests.sh t24 conv model
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/bench_profile_r1h.json > gpurun_out/bench_r1h.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r1h.log | cut -c1-200
