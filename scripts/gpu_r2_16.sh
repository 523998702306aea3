#!/bin/bash
# two-bits-ahead split decoder: parity subset, then timing on 1 GiB / 1 MiB blocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ahead" > gpurun_out/pytest_ahead.log 2>&1; tail -3 gpurun_out/pytest_ahead.log
timeout 900 bash scripts/variants.sh many 1024 1024 many:0:0:2 many:0:0:5 many:0:0:6 many:0:4:6 > gpurun_out/ahead_variants.txt 2>&1; cat gpurun_out/ahead_variants.txt
