#!/bin/bash
# round 2, first call: the GPU suite in every kernel family, then the many-block kernels on 1 GiB
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python scripts/exp_many.py 1024 1024 team:0:0 many:0:0 many:4:1 many:8:2 many:2:7 warp:0:0 > gpurun_out/exp_many_1m.txt 2>&1; echo "exp rc=$?"
cat gpurun_out/exp_many_1m.txt
timeout 600 python scripts/exp_many.py 1024 256 many:0:0 many:0:8 many:4:2 warp:0:0 > gpurun_out/exp_many_256k.txt 2>&1; echo "exp rc=$?"
cat gpurun_out/exp_many_256k.txt
timeout 300 python scripts/exp_many.py 64 1024 team:0:0 many:0:0 > gpurun_out/exp_many_64.txt 2>&1
cat gpurun_out/exp_many_64.txt
