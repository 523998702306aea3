"""Experiment: in-kernel wait / chain counters of the two-bits-ahead split decoder (a library built with
-DW3_AHEAD_PROFILE=1, see scripts/variants.sh).  64 KiB blocks so that 32-bit clocks do not wrap.
usage: W3_LIB=tools/variants/prof.so python scripts/exp_ahead_prof.py [form] [n_blocks] [group]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from corpus import w3_corpus
from weath3rb0i_b200 import capi

form = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
group = int(sys.argv[3]) if len(sys.argv) > 3 else 0
block = 64 << 10
n = n_blocks * block
ctx = capi.Context([0])
ctx.set_kernels("many")
data = w3_corpus.generate(n)
blob = ctx.compress(data, block)
ctx.set_option(capi.W3_OPT_LANE_FORM, form)
ctx.set_option(capi.W3_OPT_LANE_WARPS, group)
assert ctx.decompress(blob) == data
out = (C.c_ulonglong * 16)()
rc = ctx.lib.w3_debug_ahead_prof(out)
v = list(out)
steps = v[1]
if form >= 7:
    tl = (C.c_uint * (4 * 256))()
    ctx.lib.w3_debug_ahead_tl(tl)
    t = [list(tl[i * 256:(i + 1) * 256]) for i in range(4)]
    base = t[0][8]
    print("q: bit q decoded | model saw count q-2 | model published cand(q)   (cycles, relative)")
    for q in range(8, 72):
        print(f"{q:3d}: {(t[0][q] - base) & 0xFFFFFFFF:6d} | {(t[1][q] - base) & 0xFFFFFFFF:6d} | {(t[2][q] - base) & 0xFFFFFFFF:6d}   seen-after-decode(q-3) {(t[1][q] - t[0][q - 3]) & 0xFFFFFFFF:5d}  chain {(t[2][q] - t[1][q]) & 0xFFFFFFFF:5d}  decode-after-publish {(t[0][q] - t[2][q]) & 0xFFFFFFFF:5d}  waited for count q-3 {t[3][q] & 0xFFFF:5d}, prepare began {(t[3][q] >> 16) & 0xFFFF:5d} after the coder's stamp for count q-3")
    print(f"form {form} blocks {n_blocks} group {group}: coder {v[0] / steps:.1f} cycles/bit, {v[2] / steps:.3f} attempts/bit, {v[0] / max(v[2], 1):.1f} cycles/attempt | "
          f"model W=1: {v[4] / max(v[5], 1):.1f} cycles/activation, wait {v[7] / max(v[5], 1):.1f}, chain {v[8] / max(v[5], 1):.1f} cycles/act (rc {rc})")
    sys.exit(0)
print(f"form {form} blocks {n_blocks} group {group}: coder {v[0] / steps:.1f} cycles/step, waits that spun {v[2] / steps:.3f}/step, spun {v[3] / steps:.1f} cycles/step | "
      f"model W=1: {v[4] / max(v[5], 1):.1f} cycles/activation, spun {v[6] / max(v[5], 1):.3f}/act, {v[7] / max(v[5], 1):.1f} cycles/act, chain {v[8] / max(v[5], 1):.1f} cycles/act (rc {rc})")
