#!/bin/bash
mkdir -p gpurun_out
for v in base; do echo "== $v"; W3_LIB=$PWD/tools/variants/$v.so timeout 300 python scripts/exp_many.py 1024 1024 many:0:0:8 many:0:0:7 2>&1 | tail -n +2; done > gpurun_out/lt_variants6.txt 2>&1; cat gpurun_out/lt_variants6.txt
W3_LIB=$PWD/tools/variants/prof.so timeout 200 python scripts/exp_ahead_prof.py 8 148 1 > gpurun_out/ahead_tl3.txt 2>&1; tail -16 gpurun_out/ahead_tl3.txt
W3_LIB=$PWD/tools/variants/prof.so timeout 200 python scripts/exp_ahead_prof.py 8 1024 0 2>&1 | tail -1
