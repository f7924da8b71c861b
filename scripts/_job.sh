timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_teacher_forced.py -m gpu -q 2>&1 | tail -2
python scripts/kernel_times.py 2>&1 | grep "cla_\|decode\|per step"
python scripts/kernel_times.py skyeye_s 640 32 2>&1 | grep "decode\|per step"
