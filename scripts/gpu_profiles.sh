#!/bin/bash
# round-end evidence: launch list of the bench command, full captures of both codec kernels
# on the bench workload itself (64 MiB, 1 MiB blocks), clocks during the run
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --no-cpu"
python bench.py $ARGS > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python bench.py $ARGS > gpurun_out/prof_plain2.json 2>> gpurun_out/prof_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:w3_encode_blocks_team -s 3 -c 1 -f -o gpurun_out/prof_encode python bench.py $ARGS > gpurun_out/ncu_enc.log 2>&1
echo "enc rc=$?"
python bench.py $ARGS > gpurun_out/prof_plain3.json 2>> gpurun_out/prof_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:w3_decode_blocks_team -s 3 -c 1 -f -o gpurun_out/prof_decode python bench.py $ARGS > gpurun_out/ncu_dec.log 2>&1
echo "dec rc=$?"
ls -la gpurun_out | tail -8
