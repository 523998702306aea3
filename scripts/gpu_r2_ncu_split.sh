#!/bin/bash
# launch list of the bench command and ncu --set full of the split decoder, on the bench workload
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --no-cpu --no-config2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:w3_decode_blocks_split -s 3 -c 1 -f -o gpurun_out/r02_prof_decode_split python bench.py $ARGS > gpurun_out/ncu_dec_split.log 2>&1; echo "ncu split rc=$?"
