#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python scripts/exp_many.py 1024 1024 many:0:0:0 many:0:4:1 many:0:2:2 many:0:1:2 many:0:7:2 many:0:8:1 > gpurun_out/exp_many_1m.txt 2>&1; echo "exp rc=$?"
cat gpurun_out/exp_many_1m.txt
timeout 600 python scripts/exp_many.py 1024 256 many:0:0:0 many:0:4:2 many:0:8:1 many:0:8:2 > gpurun_out/exp_many_256k.txt 2>&1; echo "exp rc=$?"
cat gpurun_out/exp_many_256k.txt
timeout 300 python scripts/exp_many.py 1024 4096 many:0:0:0 team:0:0 many:0:2:2 > gpurun_out/exp_many_4m.txt 2>&1
cat gpurun_out/exp_many_4m.txt
