#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --mib 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/big_auto.json 2> gpurun_out/big_auto.err; echo "auto rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/big_auto.json')); print('1GiB auto', d['value'], d['compress_mbs'], d['decompress_mbs'], d['kernel_ms'], d['e2e']['value'])"
tail -2 gpurun_out/big_auto.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('64MiB', d['value'], d['compress_mbs'], d['decompress_mbs'], d['cycles_per_bit_per_warp'])"
