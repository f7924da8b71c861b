#!/bin/bash
# bench.py at N = 1, 2, 4, 8 on ONE box (the driver's SCALE run does the same at round end): whole-job images/s (weak scaling by image)
# and the config-4 tiled leg (strong scaling over 16 fixed 4K frames, NCCL all_gather + merge), with the cross-N detection checksum.
mkdir -p gpurun_out
TAG=${1:-r2g}
for n in ${2:-1 2 4 8}; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/${TAG}_scale_n$n.json 2> gpurun_out/${TAG}_scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/${TAG}_scale_n$n.json 2> gpurun_out/${TAG}_scale_n$n.err
  fi
  python - $n gpurun_out/${TAG}_scale_n$n.json <<'PY'
import json, sys
n, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    t = d["tiled4k"]
    print(f"N={n}: {d['value']:.1f} images/s (e2e {d['e2e']['value']:.1f}) | tiled {t['frames_per_s']:.2f} frames/s, {t['ms_per_step']:.2f} ms/step vs {t['untiled_ms_for_the_same_tiles']:.2f} untiled, "
          f"gather+merge {t['ms_gather_merge']:.2f} ms, sha {t['sha256_16']}, detections {t['detections']}")
except Exception as e:
    print(f"N={n}: FAILED {e}")
PY
done
