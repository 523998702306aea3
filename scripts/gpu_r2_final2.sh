#!/bin/bash
# closing run after the split decoder's last change (6 thresholds per lookup): whole GPU suite, bench, reference arm, smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final2.log 2>&1; tail -3 gpurun_out/pytest_gpu_final2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_final2_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final2_ref.json 2> gpurun_out/bench_final2_ref.err; echo "ref rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
