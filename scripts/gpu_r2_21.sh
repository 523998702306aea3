#!/bin/bash
mkdir -p gpurun_out
for v in base pull2 pull2late2; do echo "== $v"; W3_LIB=$PWD/tools/variants/$v.so timeout 300 python scripts/exp_many.py 1024 1024 many:0:0:8 2>&1 | tail -n +2; done > gpurun_out/lt_variants4.txt 2>&1; cat gpurun_out/lt_variants4.txt
export W3_LIB=$PWD/tools/variants/prof.so
for cfg in "8 1024 0" "8 148 1"; do timeout 200 python scripts/exp_ahead_prof.py $cfg; done > gpurun_out/ahead_prof3.txt 2>&1
cat gpurun_out/ahead_prof3.txt
