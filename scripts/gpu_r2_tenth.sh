#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
cat gpurun_out/bench_r2.json | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo "ref rc=$?"
timeout 700 python scripts/exp_hbm_paths.py 64 > gpurun_out/hbm_paths.jsonl 2> gpurun_out/hbm_paths.err; echo "hbm rc=$?"; cat gpurun_out/hbm_paths.jsonl; tail -3 gpurun_out/hbm_paths.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
