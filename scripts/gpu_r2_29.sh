#!/bin/bash
mkdir -p gpurun_out
for v in base fix6 fix8; do echo "== $v"; W3_LIB=$PWD/tools/variants/$v.so timeout 300 python scripts/exp_many.py 1024 1024 many:0:0:0 2>&1 | tail -n +2; done > gpurun_out/split_fix_variants.txt 2>&1; cat gpurun_out/split_fix_variants.txt
