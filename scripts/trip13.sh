mkdir -p gpurun_out
bash scripts/gpu_tests.sh t13 ops model
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/bench_profile_r1d.json > gpurun_out/bench_r1d.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r1d.log | cut -c1-300
python scripts/bench_layers.py --once --only attn_p3 > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"flash_attn" -c 1 -o gpurun_out/prof_attn2 python scripts/bench_layers.py --once --only attn_p3 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
