#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_geometry.py -x -q > gpurun_out/pytest_geo.log 2>&1; tail -3 gpurun_out/pytest_geo.log
timeout 900 python scripts/exp_hbm_1gib.py 1024 > gpurun_out/hbm_1gib.jsonl 2> gpurun_out/hbm_1gib.err; echo "hbm1g rc=$?"; cat gpurun_out/hbm_1gib.jsonl; tail -2 gpurun_out/hbm_1gib.err
( time timeout 300 python -m weath3rb0i_b200.sweep eh-ac /tmp/book1like ) > gpurun_out/sweep_book1like.txt 2>&1 || true
python -c "
from corpus import w3_corpus
open('/tmp/book1like','wb').write(w3_corpus.generate(768771))
"
( time timeout 300 python -m weath3rb0i_b200.sweep eh-ac /tmp/book1like ) > gpurun_out/sweep_book1like.txt 2>&1; tail -4 gpurun_out/sweep_book1like.txt
