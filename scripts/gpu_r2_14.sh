#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "overflowed" > gpurun_out/pytest_ovf.log 2>&1; tail -3 gpurun_out/pytest_ovf.log
timeout 600 python -m pytest tests/test_gpu_cm.py tests/test_gpu_cli.py -x -q > gpurun_out/pytest_cm.log 2>&1; tail -3 gpurun_out/pytest_cm.log
timeout 600 python scripts/exp_hbm_1gib.py 1024 > gpurun_out/hbm_1gib_v3.jsonl 2> gpurun_out/hbm_1gib_v3.err; cat gpurun_out/hbm_1gib_v3.jsonl; tail -2 gpurun_out/hbm_1gib_v3.err
