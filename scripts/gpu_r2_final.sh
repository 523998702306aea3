#!/bin/bash
# round-2 closing run on one B200: the whole GPU suite, the bench line, the reference arm, smoke(), the policy
# boundary of the decoder (4096 x 256 KiB), then the launch list and an ncu --set full of the split decoder.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_final_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_final_ref.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python scripts/exp_many.py 1024 256 many:0:0:0 many:0:0:2 many:0:0:4 > gpurun_out/policy_256k.txt 2>&1; cat gpurun_out/policy_256k.txt
timeout 300 python scripts/exp_many.py 1024 1024 many:0:0:0 many:0:0:2 > gpurun_out/policy_1m.txt 2>&1; cat gpurun_out/policy_1m.txt
ARGS="--steps 2 --warmup 3 --no-cpu --no-config2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:w3_decode_blocks_split -s 3 -c 1 -f -o gpurun_out/r02_prof_decode_split python bench.py $ARGS > gpurun_out/ncu_dec_split.log 2>&1; echo "ncu split rc=$?"
ls -la gpurun_out/r02_prof_decode_split.ncu-rep
