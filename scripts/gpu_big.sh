#!/bin/bash
mkdir -p gpurun_out
for mode in warp team; do
  W3_MODE=$mode timeout 900 python bench.py --mib 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/big_$mode.json 2> gpurun_out/big_$mode.err; echo "$mode rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/big_$mode.json")); print("$mode", d["value"], d["compress_mbs"], d["decompress_mbs"], d["kernel_ms"], d["e2e"]["value"])
except Exception as e: print("fail", e); print(open("gpurun_out/big_$mode.err").read()[-600:])
PY
done
for bs in 262144 4194304; do
  timeout 900 python bench.py --mib 1024 --block-size $bs --steps 2 --warmup 1 --no-cpu > gpurun_out/big_bs$bs.json 2> gpurun_out/big_bs$bs.err; echo "bs $bs rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/big_bs$bs.json")); print("bs$bs", d["value"], d["compress_mbs"], d["decompress_mbs"], d["kernel_ms"], d["bits_per_byte"])
except Exception as e: print("fail", e); print(open("gpurun_out/big_bs$bs.err").read()[-600:])
PY
done
