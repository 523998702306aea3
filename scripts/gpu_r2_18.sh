#!/bin/bash
# elastic split decoder: parity subset, then timing on 1 GiB / 1 MiB blocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lt4 or lt8" > gpurun_out/pytest_lt.log 2>&1; tail -3 gpurun_out/pytest_lt.log
timeout 900 bash scripts/variants.sh many 1024 1024 many:0:0:2 many:0:0:7 many:0:0:8 many:0:4:8 > gpurun_out/lt_variants.txt 2>&1; cat gpurun_out/lt_variants.txt
