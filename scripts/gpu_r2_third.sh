#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_configs.py tests/test_gpu_tierb.py -x -q -s > gpurun_out/pytest_configs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_configs.log
tail -15 gpurun_out/pytest_configs.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
cat gpurun_out/bench_r2.json; tail -5 gpurun_out/bench_r2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_r2_ref.json
python -c "import __graft_entry__ as g; g.smoke()"
nproc; free -g | head -2
