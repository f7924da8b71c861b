mkdir -p gpurun_out
bash scripts/gpu_tests.sh t20 conv ops nms model
timeout 900 python bench.py --steps 10 --warmup 3 --profile-json gpurun_out/bench_profile_r1g.json > gpurun_out/bench_r1g.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r1g.log | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
python scripts/bench_layers.py --once --only c1x1_128_128_320,c3x3_128_160,c3x3_256_80,attn_p3 > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|conv3x3_halo|flash_attn" -c 4 -o gpurun_out/prof_r1g python scripts/bench_layers.py --once --only c1x1_128_128_320,c3x3_128_160,c3x3_256_80,attn_p3 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
