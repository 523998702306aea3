"""bits/byte of the reference algorithm on the first MIB MiB of the synthetic text coded as ONE block
(the reference's own mode: one stream per file) — the yardstick of BASELINE.json configs[3]
('bits/byte loss vs single-block oracle, on the 64 MiB prefix if 1 GiB is too slow')."""
import ctypes as C, json, sys, time
sys.path.insert(0, ".")
from corpus import w3_corpus
from tests import _libs
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = mib << 20
buf = C.create_string_buffer(n)
w3_corpus.generate_into(C.addressof(buf), n)
o = _libs.load_oracle()
cap = n
out = C.create_string_buffer(cap)
ol = C.c_size_t()
t = time.time()
rc = o.w3o_compress_block(buf, n, out, cap, C.byref(ol))
assert rc == 0, rc
print(json.dumps({"prefix_mib": mib, "stream_bytes": ol.value, "bits_per_byte": round(8 * ol.value / n, 5), "seconds": round(time.time() - t, 1)}))
