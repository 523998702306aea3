#!/bin/bash
mkdir -p gpurun_out
W3_LIB=$PWD/tools/variants/prof.so timeout 200 python scripts/exp_ahead_prof.py 8 148 1 > gpurun_out/ahead_tl4.txt 2>&1; tail -14 gpurun_out/ahead_tl4.txt | cut -c1-250
