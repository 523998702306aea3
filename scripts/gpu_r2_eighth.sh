#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
tail -5 gpurun_out/pytest_parity.log
timeout 900 python scripts/exp_many.py 1024 1024 many:0:0:0:0 many:0:0:0:1 many:4:0:0:0 many:8:0:0:0 > gpurun_out/exp_enc_1m.txt 2>&1; cat gpurun_out/exp_enc_1m.txt
timeout 600 python scripts/exp_many.py 512 1024 many:0:0:0:0 many:0:0:0:1 > gpurun_out/exp_enc_512.txt 2>&1; cat gpurun_out/exp_enc_512.txt
timeout 600 python scripts/exp_many.py 256 1024 many:0:0:0:0 many:0:0:0:1 team:0:0:0:0 > gpurun_out/exp_enc_256.txt 2>&1; cat gpurun_out/exp_enc_256.txt
timeout 600 python scripts/exp_many.py 1024 256 many:0:0:0:0 many:0:0:0:1 > gpurun_out/exp_enc_256k.txt 2>&1; cat gpurun_out/exp_enc_256k.txt
