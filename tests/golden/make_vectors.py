"""Regenerates tests/golden/tier_a_vectors.json.

The reference is Rust and cannot be compiled or imported in this image, so
these vectors are NOT reference-issued: they are produced by the Python
restatement (oracle/w3_oracle.py) and were cross-checked at survey time against
an independent C++ restatement (SURVEY.md Appendix B lists the same values).
They freeze the oracle's behaviour so that a later edit cannot silently change
it.  Run from the repo root:  python tests/golden/make_vectors.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import w3_oracle as pyo  # noqa: E402

INPUTS = [
    ("empty", b""),
    ("a", b"a"),
    ("ab", b"ab"),
    ("hello", b"hello world\n"),
    ("fox", b"The quick brown fox jumps over the lazy dog.\n"),
    ("bytes_00_ff", bytes(range(256))),
    ("zeros_64", bytes(64)),
    ("ones_64", b"\xff" * 64),
]
HASHES = [
    (0, 1), (1, 1), (0x61, 8), (0x6162, 16), (0x746865, 24),
    (0x68656C6C6F20776F, 64), (0x2020202020202020, 64), (0xFFFFFFFFFFFFFFFF, 64),
    (0, 0), (0x8000000000000000, 64), (0x5555555555555555, 3), (0xAAAAAAAAAAAAAAAA, 5),
]


def main():
    out = {"streams": [], "achash": [], "ctx_trace_a": {}}
    for name, data in INPUTS:
        s = pyo.compress_block(data)
        assert pyo.decompress_block(s, len(data)) == data
        out["streams"].append(
            {"name": name, "input_hex": data.hex(), "stream_hex": s.hex(), "file_len": 12 + len(s)}
        )
    for bits, pos in HASHES:
        out["achash"].append([hex(bits), pos, pyo.achash(bits, pos)])
    tr, last = pyo.ctx_trace(b"a")
    out["ctx_trace_a"] = {"before_each_bit": tr, "after": last}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tier_a_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
