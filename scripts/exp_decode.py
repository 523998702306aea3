"""Experiment harness: time the encode and decode kernels alone (device-resident, 64 x 1 MiB blocks) and say
whether the output matched — unlike bench.py it does not stop on a mismatch, so that deliberately
unsafe kernel variants (scripts/variants.sh, W3_LIB=...) can be timed to bound an idea's worth.
Never a bench number."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from corpus import w3_corpus
from weath3rb0i_b200 import capi

MIB = 1 << 20
n, block = 64 * MIB, MIB
n_blocks = n // block
ctx = capi.Context([0])
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
w3_corpus.generate_into(h_in.data_ptr(), n)
d_in = torch.empty(n + 16, dtype=torch.uint8, device="cuda"); d_in[:n].copy_(h_in)
stride = (block + block // 8 + 64 + 15) // 16 * 16
d_scratch = torch.empty(stride * n_blocks, dtype=torch.uint8, device="cuda")
d_lens = torch.zeros(n_blocks, dtype=torch.int32, device="cuda")
d_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
d_offs = torch.zeros(n_blocks + 1, dtype=torch.int64, device="cuda")
d_packed = torch.zeros(n + n // 8 + 64, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
ctx.encode_device(d_in.data_ptr(), n, block, d_scratch.data_ptr(), stride, d_lens.data_ptr(), d_flag.data_ptr(), s)
ctx.pack_device(d_scratch.data_ptr(), stride, d_lens.data_ptr(), n_blocks, d_offs.data_ptr(), d_packed.data_ptr(), s)
def dec():
    ctx.decode_device(d_packed.data_ptr(), d_offs.data_ptr(), n_blocks, block, n, d_out.data_ptr(), d_flag.data_ptr(), s)
for _ in range(2): dec()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); reps = 3
for _ in range(reps): dec()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
mhz = 1965.0
def enc():
    ctx.encode_device(d_in.data_ptr(), n, block, d_scratch.data_ptr(), stride, d_lens.data_ptr(), d_flag.data_ptr(), s)
enc(); torch.cuda.synchronize()
a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a2.record()
for _ in range(reps): enc()
b2.record(); torch.cuda.synchronize()
ems = a2.elapsed_time(b2) / reps
print(f"{os.environ.get('W3_LIB', 'default').split('/')[-1]:16s} encode {ems:8.2f} ms  {n / 1e6 / (ems / 1e3):7.2f} MB/s  ~{ems * 1e-3 * mhz * 1e6 / (block * 8):6.1f} cycles/bit @1965MHz", flush=True)
print(f"{os.environ.get('W3_LIB', 'default').split('/')[-1]:16s} decode {ms:8.2f} ms  {n / 1e6 / (ms / 1e3):7.2f} MB/s  ~{ms * 1e-3 * mhz * 1e6 / (block * 8):6.1f} cycles/bit @1965MHz  "
      f"flag {int(d_flag.item())}  {'match' if torch.equal(d_out[:n], d_in[:n]) else 'MISMATCH'}", flush=True)
