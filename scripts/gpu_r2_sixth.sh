#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
tail -5 gpurun_out/pytest_parity.log
timeout 900 python scripts/exp_many.py 1024 1024 many:0:0:0 many:0:0:2 many:0:0:3 many:0:2:3 many:0:1:3 many:0:7:3 > gpurun_out/exp_many_1m.txt 2>&1; cat gpurun_out/exp_many_1m.txt
timeout 600 python scripts/exp_many.py 512 1024 many:0:0:0 many:0:0:2 > gpurun_out/exp_many_512.txt 2>&1; cat gpurun_out/exp_many_512.txt
timeout 600 python scripts/exp_many.py 1024 256 many:0:0:0 many:0:0:3 many:0:0:2 > gpurun_out/exp_many_256k.txt 2>&1; cat gpurun_out/exp_many_256k.txt
