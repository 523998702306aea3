#!/bin/bash
# round-2 evidence on the bench workload (1 GiB, 1 MiB blocks, one B200): launch list of the bench
# command, ncu --set full of the dominant kernels, of the gather probe, and of the HBM-resident paths
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --no-cpu --no-config2"
python bench.py $ARGS > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:w3_decode_blocks_pair -s 3 -c 1 -f -o gpurun_out/r02_prof_decode_pair python bench.py $ARGS > gpurun_out/ncu_dec.log 2>&1
echo "dec rc=$?"
ncu --set full --clock-control none --import-source on -k regex:w3_encode_blocks_team -s 3 -c 1 -f -o gpurun_out/r02_prof_encode_team python bench.py $ARGS > gpurun_out/ncu_enc.log 2>&1
echo "enc rc=$?"
ncu --set full --clock-control none -k regex:w3_k_gather_probe -s 1 -c 1 -f -o gpurun_out/r02_prof_gather_probe python bench.py $ARGS > gpurun_out/ncu_probe.log 2>&1
echo "probe rc=$?"
cat > /tmp/hbm_paths.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from corpus import w3_corpus
from weath3rb0i_b200 import capi
ctx = capi.Context([0])
data = w3_corpus.generate(64 << 20)
ctx.set_geometry(20, 3, 17)
blob = ctx.compress(data, 1 << 20)
assert ctx.decompress(blob[:]) == data if len(sys.argv) > 1 else True
ctx.set_geometry(11, 3, 8)
c = ctx.cm_compress(data[: 16 << 20], 1 << 20, 16)
print("geo", len(blob), "cm", len(c))
PY
python /tmp/hbm_paths.py > gpurun_out/hbm_paths_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"w3_k_geo_encode|w3_k_geo_hash|w3_cm_blocks" -c 3 -f -o gpurun_out/r02_prof_hbm_paths python /tmp/hbm_paths.py > gpurun_out/ncu_hbm.log 2>&1
echo "hbm rc=$?"
ls -la gpurun_out/*.ncu-rep
