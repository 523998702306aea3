import os, sys
sys.path.insert(0, os.getcwd())
from weath3rb0i_b200 import capi
from corpus import w3_corpus
ctx = capi.Context([0])
ok = True
for mode in ("team", "warp"):
    os.environ["W3_MODE"] = mode
    for n, bs in ((0, 4096), (1, 4096), (4097, 4096), (70001, 16384), (3 * 65536 + 5, 65536)):
        data = w3_corpus.generate(n, offset=4242)
        blob = ctx.compress(data, bs)
        back = ctx.decompress(blob)
        good = back == data
        ok &= good
        print(mode, n, bs, len(blob), "ok" if good else "MISMATCH", flush=True)
    leg = ctx.compress(data[:5000], legacy=True)
    ok &= ctx.decompress(leg) == data[:5000]
print("hash", ctx.achash([0x6162, 0xffffffffffffffff], [16, 64]))
ctx.close()
sys.exit(0 if ok else 1)
