import os, sys
sys.path.insert(0, os.getcwd())
os.environ["W3_MODE"]="team"
from weath3rb0i_b200 import capi
from corpus import w3_corpus
ctx = capi.Context([0])
for data in (b"a", b"ab", b"hello world\n", w3_corpus.generate(1000)):
    blob = ctx.compress(data, legacy=True)
    try:
        back = ctx.decompress(blob)
        print(len(data), "ok" if back == data else "MISMATCH", flush=True)
    except Exception as e:
        print(len(data), "ERR", e, flush=True)
