#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lt8" > gpurun_out/pytest_lt8.log 2>&1; tail -2 gpurun_out/pytest_lt8.log
for v in base sts; do echo "== $v"; W3_LIB=$PWD/tools/variants/$v.so timeout 300 python scripts/exp_many.py 1024 1024 many:0:0:8 many:0:0:7 2>&1 | tail -n +2; done > gpurun_out/lt_variants5.txt 2>&1; cat gpurun_out/lt_variants5.txt
W3_LIB=$PWD/tools/variants/prof.so timeout 200 python scripts/exp_ahead_prof.py 8 148 1 > gpurun_out/ahead_tl2.txt 2>&1; tail -25 gpurun_out/ahead_tl2.txt
W3_LIB=$PWD/tools/variants/prof.so timeout 200 python scripts/exp_ahead_prof.py 8 1024 0 2>&1 | tail -1
