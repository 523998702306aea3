#!/bin/bash
mkdir -p gpurun_out
for v in prof prof2; do echo "== $v"; W3_LIB=$PWD/tools/variants/$v.so timeout 200 python scripts/exp_ahead_prof.py 8 148 1; done > gpurun_out/ahead_tl.txt 2>&1
cat gpurun_out/ahead_tl.txt
