#!/bin/bash
# ncu pass: launch list of the bench command + one full capture per codec kernel
# (small workload: 64 blocks of 128 KiB = the same 64 warps as the bench config, 8x shorter)
mkdir -p gpurun_out
ARGS="--mib 8 --block-size 131072 --steps 2 --warmup 3 --no-cpu"
python bench.py $ARGS > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python bench.py $ARGS > gpurun_out/ncu_plain2.json 2>> gpurun_out/ncu_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:w3_encode_blocks_team -s 3 -c 1 -f -o gpurun_out/prof_encode python bench.py $ARGS > gpurun_out/ncu_enc.log 2>&1
echo "enc rc=$?"
python bench.py $ARGS > gpurun_out/ncu_plain3.json 2>> gpurun_out/ncu_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:w3_decode_blocks_team -s 3 -c 1 -f -o gpurun_out/prof_decode python bench.py $ARGS > gpurun_out/ncu_dec.log 2>&1
echo "dec rc=$?"
ls -la gpurun_out
