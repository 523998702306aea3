#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "many" > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
tail -4 gpurun_out/pytest_parity.log
timeout 900 python scripts/exp_many.py 1024 1024 many:0:0:0:0 many:8:0:0:0 many:6:0:0:0 > gpurun_out/exp_enc_1m_v2.txt 2>&1; cat gpurun_out/exp_enc_1m_v2.txt
timeout 900 python scripts/exp_many.py 1024 256 many:0:0:0:0 > gpurun_out/exp_enc_256k_v2.txt 2>&1; cat gpurun_out/exp_enc_256k_v2.txt
timeout 1200 python scripts/config4_sweep.py 3.55797 > gpurun_out/config4_sweep.jsonl 2> gpurun_out/config4_sweep.err; echo "sweep rc=$?"; cat gpurun_out/config4_sweep.jsonl
bash scripts/gpu_r2_profiles.sh
