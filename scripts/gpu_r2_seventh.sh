#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cm.py -x -q > gpurun_out/pytest_cm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cm.log
tail -15 gpurun_out/pytest_cm.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_cm.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
cat gpurun_out/bench_r2.json | cut -c1-1200
