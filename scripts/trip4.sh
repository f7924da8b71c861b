mkdir -p gpurun_out
python scripts/bench_layers.py --json gpurun_out/layers_r1a.json > gpurun_out/layers_r1a.log 2>&1; cat gpurun_out/layers_r1a.log
python scripts/bench_layers.py --once --only c1x1_128_128_320,c3x3_128_160,attn_p3 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|flash_attn" -c 3 -o gpurun_out/prof_r1a python scripts/bench_layers.py --once --only c1x1_128_128_320,c3x3_128_160,attn_p3 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu.log
