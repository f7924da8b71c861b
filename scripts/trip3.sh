mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --profile-json gpurun_out/bench_profile.json > gpurun_out/bench1.log 2>&1; echo "bench exit $?"; tail -5 gpurun_out/bench1.log
