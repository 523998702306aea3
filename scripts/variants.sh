#!/bin/bash
# Build variants of the library with extra -D flags (here), then time each on the GPU box:
#   scripts/variants.sh build name1:"-DX=1 -DY=2" name2:"..."      (in the container)
#   gpurun -- 'bash scripts/variants.sh run'                           (on the box)
# Variant libraries live in tools/variants/ (git-ignored, travel with gpurun).
cd "$(dirname "$0")/.."
mkdir -p tools/variants gpurun_out
if [ "$1" = build ]; then
  shift; rm -f tools/variants/*.so
  for v in "$@"; do
    name=${v%%:*}; flags=${v#*:}
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall $flags \
      -shared -o tools/variants/$name.so weath3rb0i_b200/csrc/w3_kernels.cu weath3rb0i_b200/csrc/w3_tierb.cu weath3rb0i_b200/csrc/w3_geo.cu weath3rb0i_b200/csrc/w3_cm.cu weath3rb0i_b200/csrc/w3_capi.cu 2>&1 | grep -E "error" &
  done
  wait; ls tools/variants
elif [ "$1" = many ]; then
  # scripts/variants.sh many MIB BLOCK_KIB cfg...: the many-block experiment harness on every variant
  shift
  for so in tools/variants/*.so; do echo "== $(basename $so .so)"; W3_LIB=$PWD/$so timeout 300 python scripts/exp_many.py "$@" 2>&1 | tail -n +2; done
elif [ "$1" = dec ]; then
  for so in tools/variants/*.so; do W3_LIB=$PWD/$so timeout 120 python scripts/exp_decode.py 2>&1 | tail -2; done
else
  for so in tools/variants/*.so; do
    name=$(basename $so .so)
    W3_LIB=$PWD/$so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_{n}.json"))
    print(f"{n:12s} roundtrip {d['value']:7.2f}  c {d['compress_mbs']:7.2f}  d {d['decompress_mbs']:7.2f}  {d['cycles_per_bit_per_warp']}", flush=True)
except Exception as e:
    print(n, "FAILED", e)
PY
  done
fi
