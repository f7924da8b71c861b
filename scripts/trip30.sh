mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t30_all.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/t30_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --profile-json gpurun_out/bench_profile_final.json > gpurun_out/bench_final.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_final.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref_final.log | cut -c1-300
