#!/bin/bash
# split-form decoder: parity subset, then timing of variants on 1 GiB / 1 MiB blocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "split" > gpurun_out/pytest_split.log 2>&1; tail -3 gpurun_out/pytest_split.log
timeout 900 bash scripts/variants.sh many 1024 1024 many:0:0:2 many:0:0:4 many:0:4:4 > gpurun_out/split_variants.txt 2>&1; cat gpurun_out/split_variants.txt
