mkdir -p gpurun_out
python scripts/probe_numerics.py skyeye_s > gpurun_out/probe_s.log 2>&1; tail -30 gpurun_out/probe_s.log
timeout 900 python bench.py --steps 5 --warmup 3 --profile-json gpurun_out/bench_profile.json > gpurun_out/bench1.log 2>&1; echo "bench exit $?"; tail -5 gpurun_out/bench1.log
