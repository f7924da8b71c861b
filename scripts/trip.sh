mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1p.log 2>&1; tail -2 gpurun_out/smoke_r1p.log
bash scripts/gpu_tests.sh t33 conv ops nms model
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r1p.log 2> gpurun_out/bench_ref_r1p.err; tail -1 gpurun_out/bench_ref_r1p.log | cut -c1-300
python bench.py --steps 10 --warmup 3 --profile-json gpurun_out/bench_profile_r1p.json > gpurun_out/bench_r1p.log 2> gpurun_out/bench_r1p.err
tail -1 gpurun_out/bench_r1p.log | cut -c1-250
SKB_ATT_PROF=1 python scripts/attn_prof.py > gpurun_out/attn_prof_r1p.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-graph > gpurun_out/ncu_r1p.log 2>&1
echo "launch list rc $?"
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 9 -c 3 -o gpurun_out/prof_r1p_attn -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-graph > gpurun_out/ncu_r1p_attn.log 2>&1
echo "attn capture rc $?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 3 -o gpurun_out/prof_r1p_halo -f python scripts/bench_layers.py --once --only c3x3_128_160,c3x3_256_80,c3x3_64_320 > gpurun_out/ncu_r1p_halo.log 2>&1
echo "halo capture rc $?"
ls -la gpurun_out/*.ncu-rep | tail -3
