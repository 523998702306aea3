#!/bin/bash
# team vs warp kernels between 148 and 1024 blocks per GPU (1 GiB in 4 MiB / 2 MiB / 1.5 MiB blocks)
mkdir -p gpurun_out
for bs in 4194304 2097152 1572864; do
  for mode in warp team; do
    W3_MODE=$mode timeout 900 python bench.py --mib 1024 --block-size $bs --steps 1 --warmup 1 --no-cpu > gpurun_out/modes_${bs}_$mode.json 2> gpurun_out/modes_${bs}_$mode.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/modes_${bs}_$mode.json")); print("bs$bs $mode blocks", d["config"]["blocks_per_gpu"], "c/d MB/s", d["compress_mbs"], d["decompress_mbs"], flush=True)
except Exception as e: print("fail", e); print(open("gpurun_out/modes_${bs}_$mode.err").read()[-400:])
PY
  done
done
