bash scripts/gpu_tests.sh r2k conv ops nms model fullsize teacher_forced tiling
timeout 600 python -m pytest tests/test_windowed_head.py -m gpu -q > gpurun_out/r2k_windowed.log 2>&1; echo "windowed exit $?"; tail -2 gpurun_out/r2k_windowed.log
python bench.py --gpus 1 --steps 20 --warmup 5 --profile-json gpurun_out/r2k_bench_per_launch.json > gpurun_out/r2k_bench_line.json 2> gpurun_out/r2k_bench_err.log; echo "bench exit $?"
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2k_bench_reference_line.json 2>> gpurun_out/r2k_bench_err.log; echo "ref exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2k_bench_line.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d['clocks'])
print(d['latency_b1']); print(d['tiled4k']['frames_per_s'], d['tiled4k']['ms_per_step'], d['tiled4k']['sha256_16'])
print(d['windowed_head']['images_per_s'], d['config2_skyeye_s']['images_per_s']); print(d['parity']['max_rel'], d['parity']['nms_rows_equal'])
print({k:(v['ms_per_step']) for k,v in d['kernels'].items()})
"
tail -2 gpurun_out/r2k_bench_reference_line.json | cut -c1-400
