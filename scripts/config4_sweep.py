"""BASELINE.json configs[3]: block-size sweep 256 KiB - 16 MiB on 1 GiB of the
synthetic text on one B200: compress / decompress MB/s (device-resident, CUDA events), bits/byte,
and the loss against the reference's own mode — the 64 MiB prefix coded as ONE block by the CPU
oracle (scripts/single_block_oracle.py: 3.5559 bits/byte; the blocked figure is taken on the same
prefix).  Per-block identity with the oracle at 4 and 16 MiB blocks is tests/test_gpu_configs.py's.
Writes one JSON line per block size."""
import json
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from corpus import w3_corpus
from weath3rb0i_b200 import capi

MIB = 1 << 20
SINGLE = float(sys.argv[1]) if len(sys.argv) > 1 else None  # bits/byte of the single-block oracle on the 64 MiB prefix
n = 1024 * MIB
ctx = capi.Context([0])
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
w3_corpus.generate_into(h_in.data_ptr(), n)
d_in = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
d_in[:n].copy_(h_in)
d_out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for kib in (256, 512, 1024, 2048, 4096, 8192, 16384):
    block = kib << 10
    nb = n // block
    stride = (block + block // 8 + 64 + 15) // 16 * 16
    d_scratch = torch.empty(stride * nb, dtype=torch.uint8, device="cuda")
    d_lens = torch.zeros(nb, dtype=torch.int32, device="cuda")
    d_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_offs = torch.zeros(nb + 1, dtype=torch.int64, device="cuda")
    d_packed = torch.zeros(n + n // 8 + 64 * nb, dtype=torch.uint8, device="cuda")

    def enc():
        ctx.encode_device(d_in.data_ptr(), n, block, d_scratch.data_ptr(), stride, d_lens.data_ptr(), d_flag.data_ptr(), s)
        ctx.pack_device(d_scratch.data_ptr(), stride, d_lens.data_ptr(), nb, d_offs.data_ptr(), d_packed.data_ptr(), s)

    def dec():
        ctx.decode_device(d_packed.data_ptr(), d_offs.data_ptr(), nb, block, n, d_out.data_ptr(), d_flag.data_ptr(), s)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    ems, dms = timed(enc), timed(dec)
    assert int(d_flag.item()) == 0 and torch.equal(d_out[:n], d_in[:n]), kib
    total = int(d_offs[nb].item())
    prefix_blocks = 64 * MIB // block  # the blocks that make up the 64 MiB prefix
    prefix = int(d_offs[prefix_blocks].item())
    line = {"block_kib": kib, "blocks": nb, "compress_mbs": round(n / 1e6 / (ems / 1e3), 1), "decompress_mbs": round(n / 1e6 / (dms / 1e3), 1),
            "round_trip_mbs": round(n / 1e6 / ((ems + dms) / 1e3), 1), "bits_per_byte_1gib": round(8 * total / n, 4),
            "bits_per_byte_64mib_prefix": round(8 * prefix / (64 * MIB), 4)}
    if SINGLE:
        line["loss_vs_single_block_pct"] = round(100 * (8 * prefix / (64 * MIB) / SINGLE - 1), 3)
    print(json.dumps(line), flush=True)
    del d_scratch, d_packed
