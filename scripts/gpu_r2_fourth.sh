#!/bin/bash
# 2 GPUs: geometry + sweep tests, the in-process multi-device tests, strong-scaling bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_geometry.py -x -q > gpurun_out/pytest_geo.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_geo.log
tail -15 gpurun_out/pytest_geo.log
timeout 900 python -m pytest tests -m gpu -x -q -k "two_devices or every_device or config3" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
python -c "
from corpus import w3_corpus
open('/tmp/book1like','wb').write(w3_corpus.generate(768771))
"
( time python -m weath3rb0i_b200.sweep /tmp/book1like ) > gpurun_out/sweep_book1like.txt 2>&1; tail -8 gpurun_out/sweep_book1like.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "bench n2 rc=$?"
cat gpurun_out/bench_r2_n2.json; tail -5 gpurun_out/bench_r2_n2.err
