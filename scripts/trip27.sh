mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "bench n8 exit $?"; tail -1 gpurun_out/bench_n8.log | cut -c1-260
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/bench_tiled.py --frames 16 --steps 3 > gpurun_out/tiled_n8.log 2>&1; tail -1 gpurun_out/tiled_n8.log | cut -c1-400
python scripts/bench_tiled.py --frames 16 --steps 2 > gpurun_out/tiled_f16_n1.log 2>&1; tail -1 gpurun_out/tiled_f16_n1.log | cut -c1-400
