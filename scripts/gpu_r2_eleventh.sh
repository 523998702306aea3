#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err; echo "bench n4 rc=$?"
cut -c1-1500 gpurun_out/bench_r2_n4.json; tail -3 gpurun_out/bench_r2_n4.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/bench_r2_ref_n4.json 2> gpurun_out/bench_r2_ref_n4.err; echo "ref n4 rc=$?"
cut -c1-300 gpurun_out/bench_r2_ref_n4.json
timeout 300 python -m pytest tests/test_gpu_configs.py -x -q -k "every_device" > gpurun_out/pytest_multi4.log 2>&1; tail -3 gpurun_out/pytest_multi4.log
