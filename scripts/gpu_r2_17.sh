#!/bin/bash
mkdir -p gpurun_out
export W3_LIB=$PWD/tools/variants/prof.so
for cfg in "6 1024 0" "5 1024 0" "6 1024 4" "6 148 1" "6 296 2"; do timeout 200 python scripts/exp_ahead_prof.py $cfg; done > gpurun_out/ahead_prof.txt 2>&1
cat gpurun_out/ahead_prof.txt
