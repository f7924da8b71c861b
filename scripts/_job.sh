bash scripts/gpu_tests.sh r2m conv ops nms model fullsize teacher_forced tiling
timeout 600 python -m pytest tests/test_windowed_head.py -m gpu -q > gpurun_out/r2m_windowed.log 2>&1; echo "windowed exit $?"; tail -1 gpurun_out/r2m_windowed.log
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2m_bench_reference_line.json 2> gpurun_out/r2m_bench_err.log; echo "ref exit $?"
python bench.py --gpus 1 --steps 20 --warmup 5 --profile-json gpurun_out/r2m_bench_per_launch.json > gpurun_out/r2m_bench_line.json 2>> gpurun_out/r2m_bench_err.log; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2m_bench_line.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d['clocks'])
print(d['latency_b1']['p50_ms'], d['tiled4k']['frames_per_s'], d['tiled4k']['sha256_16'], d['windowed_head']['images_per_s'], d['config2_skyeye_s']['images_per_s'], d['parity']['nms_rows_equal'], d['cpu_baseline']['value'])
print({k:(v['ms_per_step']) for k,v in d['kernels'].items()})
"
python scripts/kernel_times.py skyeye_l 1280 16 2>&1 | tail -32 > gpurun_out/r2m_kernel_times.txt; head -3 gpurun_out/r2m_kernel_times.txt
