mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -rA --tb=short -x 2>&1 | tail -30
