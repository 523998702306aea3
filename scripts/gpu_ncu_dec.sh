#!/bin/bash
mkdir -p gpurun_out
ARGS="--mib 8 --block-size 131072 --steps 2 --warmup 3 --no-cpu"
python bench.py $ARGS > gpurun_out/ncu_plain3.json 2>> gpurun_out/ncu_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:w3_decode_blocks_team -s 3 -c 1 -f -o gpurun_out/prof_decode python bench.py $ARGS > gpurun_out/ncu_dec.log 2>&1
echo "dec rc=$?"; cat gpurun_out/ncu_plain3.json | cut -c1-900
