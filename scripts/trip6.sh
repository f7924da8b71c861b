mkdir -p gpurun_out
bash scripts/gpu_tests.sh t6 conv model
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/bench_profile2.json > gpurun_out/bench2.log 2>&1; echo "bench exit $?"; tail -2 gpurun_out/bench2.log | cut -c1-1500
python scripts/bench_layers.py --once --only c1x1_128_128_320,c3x3_128_160,attn_p3 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|flash_attn" -c 3 -o gpurun_out/prof_r1b python scripts/bench_layers.py --once --only c1x1_128_128_320,c3x3_128_160,attn_p3 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
