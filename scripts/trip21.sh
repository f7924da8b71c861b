mkdir -p gpurun_out
python scripts/bench_tiled.py --frames 4 --steps 2 > gpurun_out/tiled_n1.log 2>&1; tail -1 gpurun_out/tiled_n1.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/bench_tiled.py --frames 4 --steps 2 > gpurun_out/tiled_n2.log 2>&1; tail -1 gpurun_out/tiled_n2.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 exit $?"; tail -1 gpurun_out/bench_n2.log | cut -c1-300
