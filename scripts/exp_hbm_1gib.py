"""The HBM-resident paths at the bench's block count (1 GiB, 1024 x 1 MiB blocks, one B200), through the
C ABI with host buffers (wall clock): run-time geometry (20, 3, 17) compress, and w3cm (2^16 cells per
block) compress + decompress.  One JSON line each, with the gather-roofline arithmetic of SURVEY.md §8(d):
geometry: 8 counter sectors read + 8 written per input byte; w3cm: 2 nibbles x 6 models x (3 + 3) sectors."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import ctypes as C
from corpus import w3_corpus
from weath3rb0i_b200 import capi

n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1 << 30
buf = C.create_string_buffer(n)
w3_corpus.generate_into(C.addressof(buf), n)
data = buf.raw
ctx = capi.Context([0])
rate = ctx.sector_gather_rate(4 << 30)
t = time.perf_counter(); blob = ctx.cm_compress(data, 1 << 20, 16); tc = time.perf_counter() - t
t = time.perf_counter(); back = ctx.cm_decompress(blob); td = time.perf_counter() - t
assert back == data
for name, sec in (("compress", tc), ("decompress", td)):
    print(json.dumps({"path": "w3cm", "op": name, "mib": n >> 20, "blocks": n >> 20, "mbs": round(n / 1e6 / sec, 1), "bits_per_byte": round(8 * len(blob) / n, 4),
                      "sectors_per_s": round(72 * n / sec / 1e9, 2), "gather_peak_g_sectors": round(rate / 1e9, 2), "gather_frac": round(72 * n / sec / rate, 4)}), flush=True)
ctx.set_geometry(20, 3, 17)
t = time.perf_counter(); blob = ctx.compress(data, 1 << 20); tc = time.perf_counter() - t
ctx.set_geometry(11, 3, 8)
print(json.dumps({"path": "geometry (20,3,17)", "op": "compress", "mib": n >> 20, "mbs": round(n / 1e6 / tc, 1), "bits_per_byte": round(8 * len(blob) / n, 4),
                  "sectors_per_s": round(16 * n / tc / 1e9, 2), "gather_frac": round(16 * n / tc / rate, 4)}), flush=True)
