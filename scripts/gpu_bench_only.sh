#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from weath3rb0i_b200 import capi
from corpus import w3_corpus
ctx = capi.Context([0])
data = w3_corpus.generate(1 << 20, offset=777)
for mode in ("team", "warp"):
    os.environ["W3_MODE"] = mode
    blob = ctx.compress(data, 1 << 17)
    print(mode, "ok" if ctx.decompress(blob) == data else "MISMATCH", flush=True)
PY
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['compress_mbs'], d['decompress_mbs'], d['cycles_per_bit_per_warp'])"
tail -2 gpurun_out/bench.err
