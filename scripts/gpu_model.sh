#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q > gpurun_out/pytest_model.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_model.log
