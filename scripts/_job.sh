python scripts/bench_nms_stress.py > gpurun_out/r2k_nms_stress.json 2> gpurun_out/r2k_nms_stress.err; echo "stress exit $?"
bash scripts/ncu_round.sh r2k 2>&1 | tail -30
