mkdir -p gpurun_out
bash scripts/gpu_tests.sh t18 conv ops model
timeout 300 python scripts/bench_layers.py --only c3x3_128_160,c3x3_256_80,c3x3_512_40,c3x3_64_320 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/bench_profile_r1f.json > gpurun_out/bench_r1f.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r1f.log | cut -c1-300
