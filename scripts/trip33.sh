mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-latency > gpurun_out/plain_bench3.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-latency > gpurun_out/ncu_bench3.log 2>&1; echo "ncu launches exit $?"
python scripts/bench_layers.py --once --only c1x1_128_128_320,c3x3_128_160,ff0_256_1024_160,c3x3s2_128_256 > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|conv3x3_halo" -c 4 -o gpurun_out/prof_r1h python scripts/bench_layers.py --once --only c1x1_128_128_320,c3x3_128_160,ff0_256_1024_160,c3x3s2_128_256 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
