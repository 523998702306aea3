"""Throughput of the HBM-resident paths through the C ABI (host buffers, wall clock): the run-time
geometry (counters in HBM) and the hashed context-mixing stack, on MIB MiB of the synthetic text in
1 MiB blocks.  Prints one JSON line per configuration."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
from corpus import w3_corpus
from weath3rb0i_b200 import capi

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
data = w3_corpus.generate(mib << 20)
ctx = capi.Context([0])
block = 1 << 20


def timed(fn):
    fn()
    t = time.perf_counter()
    r = fn()
    return r, time.perf_counter() - t


base = ctx.compress(data, block)
for geo in ((16, 3, 13), (20, 3, 17), (24, 3, 21)):
    ctx.set_geometry(*geo)
    blob, tc = timed(lambda: ctx.compress(data, block))
    ctx.set_geometry(11, 3, 8)
    dec_mib = min(mib, 16)  # the general decoder has the literal hash on its chain
    ctx.set_geometry(*geo)
    small = ctx.compress(data[: dec_mib << 20], block)
    ctx.set_geometry(11, 3, 8)
    back, td = timed(lambda: ctx.decompress(small))
    assert back == data[: dec_mib << 20]
    print(json.dumps({"path": "geometry", "geo": geo, "mib": mib, "compress_mbs": round(len(data) / 1e6 / tc, 1),
                      "decompress_mbs": round((dec_mib << 20) / 1e6 / td, 1), "decode_mib": dec_mib,
                      "bits_per_byte": round(8 * len(blob) / len(data), 4), "vs_11_3_8": round(len(blob) / len(base), 4)}), flush=True)
for log_cells in (16, 18, 20):
    blob, tc = timed(lambda: ctx.cm_compress(data, block, log_cells))
    back, td = timed(lambda: ctx.cm_decompress(blob))
    assert back == data
    print(json.dumps({"path": "w3cm", "log_cells": log_cells, "mib": mib, "compress_mbs": round(len(data) / 1e6 / tc, 1),
                      "decompress_mbs": round(len(data) / 1e6 / td, 1), "bits_per_byte": round(8 * len(blob) / len(data), 4),
                      "vs_11_3_8": round(len(blob) / len(base), 4)}), flush=True)
