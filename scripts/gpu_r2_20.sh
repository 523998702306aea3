#!/bin/bash
mkdir -p gpurun_out
timeout 900 bash scripts/variants.sh many 1024 1024 many:0:0:6 many:0:0:8 > gpurun_out/lt_variants3.txt 2>&1; cat gpurun_out/lt_variants3.txt
