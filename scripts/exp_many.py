"""Experiment harness for the many-block regime: times the encode and the decode kernel alone
(device-resident input, CUDA events) for a list of launch options and says whether the round trip
matched and whether the packed streams equal those of the first configuration.  Never a bench number.

usage: python scripts/exp_many.py MIB BLOCK_KIB "kernels:enc_teams:lane_warps[:lane_form[:enc_coder]]" ...
"""
import hashlib
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from corpus import w3_corpus
from weath3rb0i_b200 import capi

MIB = 1 << 20
mib, block = int(sys.argv[1]), int(sys.argv[2]) << 10
configs = sys.argv[3:] or ["auto:0:0"]
n = mib * MIB
n_blocks = (n + block - 1) // block
ctx = capi.Context([0])
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
w3_corpus.generate_into(h_in.data_ptr(), n)
d_in = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
d_in[:n].copy_(h_in)
stride = (block + block // 8 + 64 + 15) // 16 * 16
d_scratch = torch.empty(stride * n_blocks, dtype=torch.uint8, device="cuda")
d_lens = torch.zeros(n_blocks, dtype=torch.int32, device="cuda")
d_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
d_offs = torch.zeros(n_blocks + 1, dtype=torch.int64, device="cuda")
d_packed = torch.zeros(n + n // 8 + 64, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
mhz = 1965.0


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


first = None
print(f"# {mib} MiB, {block >> 10} KiB blocks, {n_blocks} blocks", flush=True)
for cfg in configs:
    k, teams, warps, *rest = cfg.split(":")
    ctx.set_kernels(k)
    ctx.set_option(capi.W3_OPT_LANE_FORM, int(rest[0]) if rest else 0)
    ctx.set_option(capi.W3_OPT_ENC_CODER, int(rest[1]) if len(rest) > 1 else 0)
    ctx.set_option(capi.W3_OPT_ENC_TEAMS, int(teams))
    ctx.set_option(capi.W3_OPT_LANE_WARPS, int(warps))
    d_flag.zero_()
    enc = lambda: ctx.encode_device(d_in.data_ptr(), n, block, d_scratch.data_ptr(), stride, d_lens.data_ptr(), d_flag.data_ptr(), s)
    ems = timed(enc, 2)
    ctx.pack_device(d_scratch.data_ptr(), stride, d_lens.data_ptr(), n_blocks, d_offs.data_ptr(), d_packed.data_ptr(), s)
    torch.cuda.synchronize()
    total = int(d_offs[n_blocks].item())
    sha = hashlib.sha256(d_packed[:total].cpu().numpy().tobytes()).hexdigest()[:16]
    first = first or sha
    d_out.zero_()
    dec = lambda: ctx.decode_device(d_packed.data_ptr(), d_offs.data_ptr(), n_blocks, block, n, d_out.data_ptr(), d_flag.data_ptr(), s)
    dms = timed(dec, 2)
    ok = torch.equal(d_out[:n], d_in[:n])
    print(f"{cfg:14s} encode {ems:9.2f} ms {n / 1e6 / (ems / 1e3):8.1f} MB/s | decode {dms:9.2f} ms {n / 1e6 / (dms / 1e3):8.1f} MB/s | "
          f"SM-cycles/bit enc {ems * 1e-3 * mhz * 1e6 * 148 / (n * 8):6.1f} dec {dms * 1e-3 * mhz * 1e6 * 148 / (n * 8):6.1f} | "
          f"flag {int(d_flag.item())} {'match' if ok else 'MISMATCH'} streams {'same' if sha == first else 'DIFFER'} {total}", flush=True)
