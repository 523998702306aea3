"""Fixed-seed synthetic English-like text.

Python twin of corpus/w3_corpus.c — must produce identical bytes
(tests/test_corpus.py).  `generate()` uses the C library when it is built
(fast, threaded) and `generate_py()` is the pure-Python statement of the same
generator, used to cross-check it on small sizes.
"""
from __future__ import annotations

import ctypes
import os

SEED = 0x77333069
GOLDEN = 0x9E3779B97F4A7C15
CHUNK = 65536
VOCAB = 8192
M64 = (1 << 64) - 1
# frozen digests of the first 1 MiB / 64 MiB of the stream (also in BASELINE.md)
SHA256_1MIB = "10b1957134262774dae2cd90f01ed794b00e3a53a1ebc4fe00bc77040590b290"
SHA256_64MIB = "3ef7833dde9d2047ca41de15a63ddbc41f64238b7933fe0c4b3ee23d2fc9bca0"

FUNCTION_WORDS = (
    "the of and to a in that is was he for it with as his on be at by i this had "
    "not are but from or have an they which one you were her all she there would "
    "their we him been has when who will more no if out so said what up its about "
    "into than them can only other new"
).split()
LEN_W = (2, 14, 34, 46, 44, 38, 30, 20, 13, 8, 4, 3)
VOWELS, VOWEL_W = "eaoiuy", (330, 220, 200, 190, 70, 14)
CONS = "tnshrdlcmwfgpbvkjxqz"
CONS_W = (150, 112, 105, 101, 99, 71, 67, 46, 40, 39, 37, 33, 32, 25, 16, 13, 3, 3, 2, 30)


class _Rng:
    def __init__(self, seed):
        self.s = seed & M64

    def next(self):
        self.s = (self.s + GOLDEN) & M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)


def _pick(r, w, total):
    x = r % total
    for i, wi in enumerate(w):
        if x < wi:
            return i
        x -= wi
    return len(w) - 1


_vocab = None


def _build_vocab():
    global _vocab
    if _vocab is not None:
        return _vocab
    assert len(FUNCTION_WORDS) == 64
    rng = _Rng(SEED)
    words, cum, total = [], [], 0
    for i in range(VOCAB):
        if i < 64:
            words.append(FUNCTION_WORDS[i].encode())
        else:
            L = 1 + _pick(rng.next() >> 16, LEN_W, 256)
            L = max(L, 2)
            vowel = ((rng.next() >> 20) & 3) == 0
            w = bytearray()
            for _ in range(L):
                r = rng.next() >> 16
                if vowel:
                    w.append(ord(VOWELS[_pick(r, VOWEL_W, 1024)]))
                else:
                    w.append(ord(CONS[_pick(r, CONS_W, 1024)]))
                t = (r >> 24) & 7
                vowel = (t == 0) if vowel else (t < 6)
            words.append(bytes(w))
        total += (1 << 32) // (i + 1)
        cum.append(total)
    _vocab = (words, cum, total)
    return _vocab


def _zipf_rank(cum, total, r):
    x = (r >> 11) % total
    lo, hi = 0, VOCAB - 1
    while lo < hi:
        mid = (lo + hi) >> 1
        if x < cum[mid]:
            hi = mid
        else:
            lo = mid + 1
    return lo


def _gen_chunk(k):
    words, cum, total = _build_vocab()
    rng = _Rng(SEED ^ ((k * GOLDEN) & M64))
    buf = bytearray()
    recent = []
    sentence_start = True
    while len(buf) < CHUNK:
        r = rng.next()
        if ((r >> 8) & 63) == 0:
            nd = 1 + ((r >> 16) & 3)
            d = r >> 20
            for _ in range(nd):
                buf.append(ord("0") + d % 10)
                d //= 10
        else:
            if len(recent) == 8 and ((r >> 16) & 15) == 0:
                w = recent[(r >> 20) & 7]
            else:
                w = _zipf_rank(cum, total, rng.next())
            if len(recent) < 8:
                recent.append(w)
            else:
                recent[(r >> 24) & 7] = w
            word = words[w]
            if sentence_start:
                word = bytes([word[0] - 32]) + word[1:]
            buf += word
        sentence_start = False
        e = (r >> 32) % 168
        if e < 12:
            buf += b".\n" if ((r >> 44) % 5) == 0 else b". "
            sentence_start = True
        elif e < 26:
            buf += b", "
        else:
            buf += b" "
    return bytes(buf[:CHUNK])


def generate_py(n: int, offset: int = 0) -> bytes:
    """Bytes [offset, offset+n) of the corpus stream, pure Python."""
    if n == 0:
        return b""
    first, last = offset // CHUNK, (offset + n - 1) // CHUNK
    data = b"".join(_gen_chunk(k) for k in range(first, last + 1))
    lo = offset - first * CHUNK
    return data[lo : lo + n]


_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libw3corpus.so")
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is not built — run `python -c 'import __graft_entry__ as g; g.build()'`"
            )
        _lib = ctypes.CDLL(path)
        _lib.w3c_generate.restype = ctypes.c_int
        _lib.w3c_generate.argtypes = [
            ctypes.c_void_p,
            ctypes.c_uint64,
            ctypes.c_uint64,
            ctypes.c_int,
        ]
    return _lib


def generate_into(ptr: int, n: int, offset: int = 0, threads: int | None = None) -> None:
    """Fill host memory at address `ptr` with n corpus bytes (C generator)."""
    if threads is None:
        threads = min(os.cpu_count() or 1, 64)
    rc = _load().w3c_generate(ctypes.c_void_p(ptr), offset, n, threads)
    if rc != 0:
        raise RuntimeError(f"w3c_generate failed: {rc}")


def generate(n: int, offset: int = 0, threads: int | None = None) -> bytes:
    buf = ctypes.create_string_buffer(max(n, 1))
    generate_into(ctypes.addressof(buf), n, offset, threads)
    return buf.raw[:n]
