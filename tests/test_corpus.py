"""Synthetic corpus generator: C and Python twins agree,
prefix-stable, 7-bit ASCII, oracle ratio in the accepted band."""
import hashlib

from corpus import w3_corpus


def test_c_matches_python():
    n = 3 * 65536 + 123
    assert w3_corpus.generate(n) == w3_corpus.generate_py(n)
    assert w3_corpus.generate(1000, offset=65536 - 500) == w3_corpus.generate_py(1000, offset=65536 - 500)


def test_prefix_stable_and_ascii():
    a = w3_corpus.generate(1 << 20)
    b = w3_corpus.generate(1 << 18)
    assert a[: 1 << 18] == b
    assert w3_corpus.generate(4096, offset=(1 << 20) - 4096) == a[-4096:]
    assert max(a) < 128
    assert w3_corpus.generate(1 << 20, threads=1) == a


def test_frozen_digest_1mib():
    # frozen in BASELINE.md; a change here invalidates every recorded number
    assert hashlib.sha256(w3_corpus.generate(1 << 20)).hexdigest() == w3_corpus.SHA256_1MIB


def test_oracle_ratio_in_band(libs):
    data = w3_corpus.generate(1 << 18)
    r = len(libs.oracle.compress_block(data)) / len(data)
    assert 0.40 <= r <= 0.55, r
