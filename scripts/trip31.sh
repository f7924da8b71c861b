mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_bench2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"flash_attn" -s 9 -c 3 -o gpurun_out/prof_bench_attn python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench2.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_bench2.log | cut -c1-200
