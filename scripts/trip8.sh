mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_r1q.log 2> gpurun_out/bench_n8_r1q.err
tail -1 gpurun_out/bench_n8_r1q.log | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/bench_tiled.py --frames 16 --steps 3 > gpurun_out/tiled_n8_r1q.log 2> gpurun_out/tiled_n8_r1q.err
tail -1 gpurun_out/tiled_n8_r1q.log
python scripts/bench_tiled.py --frames 16 --steps 2 > gpurun_out/tiled_n1_r1q.log 2> gpurun_out/tiled_n1_r1q.err
tail -1 gpurun_out/tiled_n1_r1q.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_r1q.log 2> gpurun_out/bench_n1_r1q.err
tail -1 gpurun_out/bench_n1_r1q.log | cut -c1-200
