#!/bin/bash
# One gpurun call: the plain run (must exit 0), the ncu launch list of the same command, and `--set full` captures of the top kernels.
# usage: bash scripts/ncu_round.sh <tag>
TAG=${1:-r2e}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-tiled --no-variants --no-cpu-baseline --no-latency --no-graph"
$CMD > gpurun_out/${TAG}_plain.log 2> gpurun_out/${TAG}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/${TAG}_plain2.log 2> gpurun_out/${TAG}_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 3 -c 3 -f -o gpurun_out/${TAG}_attn $CMD > gpurun_out/${TAG}_ncu_attn.log 2>&1
echo "attention capture exit $?"
$CMD > gpurun_out/${TAG}_plain3.log 2> gpurun_out/${TAG}_plain3.err && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 31 -c 3 -f -o gpurun_out/${TAG}_halo $CMD > gpurun_out/${TAG}_ncu_halo.log 2>&1
echo "halo capture exit $?"
$CMD > gpurun_out/${TAG}_plain4.log 2> gpurun_out/${TAG}_plain4.err && \
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 100 -c 6 -f -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "gemm capture exit $?"
WCMD="python scripts/bench_window_attn.py"
$WCMD > gpurun_out/${TAG}_plain5.log 2> gpurun_out/${TAG}_plain5.err && \
ncu --set full --clock-control none --import-source on -k regex:window_attn_tc -s 3 -c 1 -f -o gpurun_out/${TAG}_wattn $WCMD > gpurun_out/${TAG}_ncu_wattn.log 2>&1
echo "window attention capture exit $?"
ls -la gpurun_out/ | grep ${TAG}
