#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_geometry.py -x -q > gpurun_out/pytest_geo.log 2>&1; tail -3 gpurun_out/pytest_geo.log
python -c "
from corpus import w3_corpus
open('/tmp/book1like','wb').write(w3_corpus.generate(768771))
"
( time timeout 300 python -m weath3rb0i_b200.sweep ordern /tmp/book1like ) > gpurun_out/sweep_ordern_book1like.txt 2>&1; tail -5 gpurun_out/sweep_ordern_book1like.txt
timeout 900 python scripts/exp_hbm_1gib.py 1024 > gpurun_out/hbm_1gib.jsonl 2> gpurun_out/hbm_1gib.err; echo "hbm1g rc=$?"; cat gpurun_out/hbm_1gib.jsonl; tail -2 gpurun_out/hbm_1gib.err
cat > /tmp/hbm_small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from corpus import w3_corpus
from weath3rb0i_b200 import capi
ctx = capi.Context([0])
data = w3_corpus.generate(32 << 18)
ctx.set_geometry(20, 3, 17)
blob = ctx.compress(data, 1 << 18)
ctx.set_geometry(11, 3, 8)
c = ctx.cm_compress(data, 1 << 18, 16)
print("geo", len(blob), "cm", len(c))
PY
timeout 120 python /tmp/hbm_small.py > gpurun_out/hbm_small_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:"w3_k_geo_encode|w3_cm_blocks" -c 2 -f -o gpurun_out/r02_prof_hbm_small python /tmp/hbm_small.py > gpurun_out/ncu_hbm_small.log 2>&1
echo "ncu small rc=$?"
