mkdir -p gpurun_out
bash scripts/gpu_tests.sh t31 conv model
timeout 600 python scripts/bench_layers.py --iters 7 --only c3x3_128_160,c3x3_256_80,c3x3_64_320,c3x3_512_40,c1x1_128_128_160 2>&1 | tee gpurun_out/layers_2issuers.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1n.log 2> gpurun_out/bench_r1n.err
tail -1 gpurun_out/bench_r1n.log | cut -c1-200
