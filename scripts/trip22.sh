mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "flash" --tb=short > gpurun_out/t22_attn.log 2>&1; echo "attn tests exit $?"; tail -3 gpurun_out/t22_attn.log
for P in 0 8; do echo "POLY $P"; SKB_ATT_POLY=$P timeout 300 python scripts/bench_layers.py --only attn_p3,attn_p4,attn_p5 2>&1 | tail -3; done
