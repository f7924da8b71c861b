timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "layernorm or transformer or cbam" 2>&1 | tail -2
for h in 0 1 0 1; do echo "SKB_LN_HOIST=$h"; SKB_LN_HOIST=$h python scripts/kernel_times.py 2>&1 | grep "layernorm\|cbam_mlp"; done
