import os, sys
sys.path.insert(0, os.getcwd())
os.environ["W3_MODE"]="team"
from weath3rb0i_b200 import capi
capi.LIB_PATH = os.path.join(os.getcwd(), "tools", "libw3codec_dbg.so")  # scripts/build_dbg.sh
from corpus import w3_corpus
ctx = capi.Context([0])
data = w3_corpus.generate(1 << 18, offset=12345)
blob = ctx.compress(data, 1 << 17)
back = ctx.decompress(blob)
print(len(data), "ok" if back == data else "MISMATCH", flush=True)
