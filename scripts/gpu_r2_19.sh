#!/bin/bash
mkdir -p gpurun_out
export W3_LIB=$PWD/tools/variants/prof.so
for cfg in "6 1024 0" "8 1024 0" "7 1024 0" "8 148 1"; do timeout 200 python scripts/exp_ahead_prof.py $cfg; done > gpurun_out/ahead_prof2.txt 2>&1
cat gpurun_out/ahead_prof2.txt
export W3_LIB=$PWD/tools/variants/base.so
timeout 600 python scripts/exp_many.py 1024 1024 many:0:0:2 many:0:0:6 many:0:0:8 > gpurun_out/lt_variants2.txt 2>&1; cat gpurun_out/lt_variants2.txt
