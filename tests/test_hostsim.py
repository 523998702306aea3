"""w3_core.cuh (the statements the CUDA kernels execute) compiled for the CPU
and compared with the oracle: closed-form renormalisation, threshold-table
hash, float-estimate Counter::p, word writer.  No GPU needed."""
import random

import pytest

from corpus import w3_corpus
from tests._libs import BOOK1, ENWIK7, HostSim


def test_table_builds_monotone(libs):
    assert libs.hostsim.bad == 0
    thr, sing = libs.hostsim.thresholds()
    for a in range(8):
        assert thr[a][:255] == sorted(thr[a][:255])
        assert thr[a][255] == (1 << 64) - 1


def test_hash_table_equals_oracle_random(libs):
    rnd = random.Random(11)
    for i in range(60000):
        bits = rnd.getrandbits(64)
        if i % 3 == 0:
            bits &= (1 << rnd.randrange(1, 64)) - 1
        pos = rnd.randrange(0, 1 << 24)
        assert libs.hostsim.hash(bits, pos)[0] == libs.oracle.achash(bits, pos)[0]


def _rev64(v):
    return int(f"{v:064b}"[::-1], 2)


def test_hash_table_at_every_threshold(libs):
    """Keys at, just below and just above each of the 8 x 255 thresholds —
    including the singular keys, which must take the literal-loop fallback."""
    thr, sing = libs.hostsim.thresholds()
    n_sing = 0
    for a in range(8):
        for j in range(255):
            for key in (thr[a][j] - 1, thr[a][j], thr[a][j] + 1):
                if not 0 <= key < 1 << 64:
                    continue
                bits = _rev64(key)
                got, slow = libs.hostsim.hash(bits, a)
                assert got == libs.oracle.achash(bits, a)[0], (a, j, hex(key))
                if key == thr[a][j]:
                    assert slow
            if sing[a][j]:
                n_sing += 1
                _, idx = libs.hostsim.hash_exact(_rev64(thr[a][j]), a)
                assert idx < 8
    assert n_sing > 0


def test_hash_table_other_model(libs):
    hs = HostSim(ENWIK7, 8)
    assert hs.bad == 0
    rnd = random.Random(5)
    for _ in range(20000):
        bits, pos = rnd.getrandbits(64), rnd.randrange(64)
        assert hs.hash(bits, pos)[0] == libs.oracle.achash(bits, pos, ENWIK7, 8)[0]


def test_counter_p_estimate_is_exact(libs):
    """w3_counter_p (float estimate + integer fix-up) and w3_counter_p_both (both updates of a state
    from one reciprocal, one correction each way) against the u64 statement of counter.rs:13-18;
    the device runs the same sweep over all 65535^2 states (tests/test_gpu_parity.py)."""
    lib = libs.hostsim.lib
    assert lib.hs_counter_p_sweep(0, 2048, 1) == 0
    assert lib.hs_counter_p_sweep(2048, 65535, 61) == 0
    assert lib.hs_counter_p_sweep(65400, 65535, 1) == 0


def test_counter_update(libs):
    import ctypes as C

    rnd = random.Random(3)
    for _ in range(20000):
        c0, c1, bit = rnd.choice([rnd.randrange(65535), 65534, 65533, 0]), rnd.randrange(65535), rnd.getrandbits(1)
        c = (C.c_uint16 * 2)(c0, c1)
        libs.oracle.lib.w3o_counter_update(c, bit)
        got = libs.hostsim.lib.hs_counter_update(c0 | (c1 << 16), bit)
        assert (got & 0xFFFF, got >> 16) == (c[0], c[1])


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 8, 9, 63, 64, 65, 1000, 1 << 16])
def test_block_equals_oracle_text(libs, n):
    data = w3_corpus.generate(n, offset=12345)
    c = libs.hostsim.compress_block(data)
    assert c == libs.oracle.compress_block(data)
    assert libs.hostsim.decompress_block(c, n) == data


def test_block_equals_oracle_hostile_inputs(libs):
    rnd = random.Random(9)
    cases = [
        bytes(rnd.getrandbits(8) for _ in range(30000)),
        bytes(30000),
        b"\xff" * 30000,
        bytes(range(256)) * 50,
        b"\x80" * 9000 + b"\x01" * 9000,
        bytes((i * 37) & 0xFF for i in range(20000)),
    ]
    for data in cases:
        c = libs.hostsim.compress_block(data)
        assert c == libs.oracle.compress_block(data)
        assert libs.hostsim.decompress_block(c, len(data)) == data
        # truncated stream: zero bits past the end (io.rs:23-25)
        cut = c[: len(c) // 3]
        assert libs.hostsim.decompress_block(cut, len(data)) == libs.oracle.decompress_block(cut, len(data))


def test_counter_halving_reached(libs):
    """A block long enough that some counter crosses 65535 and is halved
    (counter.rs:22-25): alignment-0 context of a constant byte."""
    data = b"e" * 70000
    c = libs.hostsim.compress_block(data)
    assert c == libs.oracle.compress_block(data)
    assert libs.hostsim.decompress_block(c, len(data)) == data


def test_rank_relabelled_counters_change_nothing(libs):
    """The kernels index the counter table by (rank << 3) | alignment, a
    bijective relabelling of (hash << 3) | alignment: same bytes out."""
    rnd = random.Random(13)
    cases = [w3_corpus.generate(50000, offset=5), bytes(rnd.getrandbits(8) for _ in range(20000)),
             b"e" * 70000, bytes(3000), b""]
    libs.hostsim.lib.hs_set_ranked(1)
    try:
        for data in cases:
            c = libs.hostsim.compress_block(data)
            assert c == libs.oracle.compress_block(data)
            assert libs.hostsim.decompress_block(c, len(data)) == data
    finally:
        libs.hostsim.lib.hs_set_ranked(0)


def test_width_carrying_coder_changes_nothing(libs):
    """w3_coder_step / w3_coder_renorm (state = (x1, width), total shift from one clz — what the
    team kernels execute) against the oracle: same streams, same decodes, truncated streams too."""
    rnd = random.Random(17)
    cases = [w3_corpus.generate(60000, offset=9), bytes(rnd.getrandbits(8) for _ in range(30000)),
             b"e" * 70000, bytes(30000), b"\xff" * 30000, bytes(range(256)) * 40, b"a", b""]
    libs.hostsim.lib.hs_set_coder(1)
    try:
        for data in cases:
            c = libs.hostsim.compress_block(data)
            assert c == libs.oracle.compress_block(data)
            assert libs.hostsim.decompress_block(c, len(data)) == data
            cut = c[: len(c) // 3]
            assert libs.hostsim.decompress_block(cut, len(data)) == libs.oracle.decompress_block(cut, len(data))
    finally:
        libs.hostsim.lib.hs_set_coder(0)


def test_renorm_forms_against_the_literal_loops(tmp_path):
    """tools/renorm_check.cpp: both closed forms of the renormalisation against the loops of
    arithmetic_coder.rs:51-62 on a few million states (the committed default run is 6e8)."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "renorm_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(root, "tools", "renorm_check.cpp")])
    out = subprocess.run([exe, "3000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches" in out.stdout


def test_fixed_latency_rank_lookup(libs):
    """w3_rank_lookup_fixed (decode speculators) agrees with the general lookup
    wherever it answers, and answers almost always."""
    import ctypes as C

    lib = libs.hostsim.lib
    lib.hs_rank_fixed.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
    lib.hs_rank.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
    rnd = random.Random(17)
    data = w3_corpus.generate(20000, offset=999)
    keys, hist, t = [], 0, 0
    for byte in data[:4000]:
        for k in range(7, -1, -1):
            hist = ((hist << 1) | ((byte >> k) & 1)) & ((1 << 64) - 1)
            t += 1
            keys.append((hist, t))
    keys += [(rnd.getrandbits(64), rnd.randrange(64)) for _ in range(30000)]
    thr, _ = libs.hostsim.thresholds()
    for a in range(8):
        for j in range(0, 255, 7):
            keys.append((int(f"{thr[a][j]:064b}"[::-1], 2), a))
    miss = 0
    for bits, pos in keys:
        f = lib.hs_rank_fixed(libs.hostsim.tab, bits, pos)
        g = lib.hs_rank(libs.hostsim.tab, bits, pos)
        if f < 0:
            miss += 1
        else:
            assert f == g, (hex(bits), pos)
    assert miss < len(keys) // 50


# ---- the lane decoder (w3_lane.cuh), on the CPU ------------------------------------------------


@pytest.mark.parametrize("pair", [False, True, "cache", "split"], ids=["lane", "pair", "pair-cached-p", "split"])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 31, 100, 4097, 70001])
def test_lane_decoder_equals_oracle(libs, n, pair):
    """The statements every lane of w3_decode_blocks_lane / _pair executes, at every start
    misalignment of a packed stream and with the next stream's bytes right behind it."""
    data = w3_corpus.generate(n, offset=4242)
    stream = libs.oracle.compress_block(data)
    for mis in range(4):
        assert libs.hostsim.lane_decode_block(stream, n, mis, pair=pair) == data, (n, mis)
        assert libs.hostsim.lane_decode_block(stream, n, mis, junk=0xFF, pair=pair) == data, (n, mis)


@pytest.mark.parametrize("pair", [False, True, "cache", "split"], ids=["lane", "pair", "pair-cached-p", "split"])
def test_lane_decoder_hostile_and_truncated(libs, pair):
    rnd = random.Random(17)
    cases = [bytes(rnd.getrandbits(8) for _ in range(20000)), bytes(30000), b"\xff" * 30000,
             b"e" * 70000, bytes(range(256)) * 40, b"\x80" * 5000 + b"\x01" * 5000]
    for data in cases:
        stream = libs.oracle.compress_block(data)
        assert libs.hostsim.lane_decode_block(stream, len(data), 1, pair=pair) == data
    # a cut stream decodes with zero bits past its end, exactly as ACReader does (io.rs:23-25)
    data = w3_corpus.generate(6000)
    stream = libs.oracle.compress_block(data)
    for cut in (0, 1, 3, 4, 5, len(stream) // 2, len(stream) - 1):
        want = libs.oracle.decompress_block(stream[:cut], len(data))
        for mis in (0, 3):
            assert libs.hostsim.lane_decode_block(stream[:cut], len(data), mis, pair=pair) == want, cut


@pytest.mark.parametrize("pair", [False, True, "cache", "split"], ids=["lane", "pair", "pair-cached-p", "split"])
def test_lane_decoder_keys_on_thresholds(libs, pair):
    """Histories that put the key exactly on a threshold (singular keys: general search) in the stream."""
    thr, _ = libs.hostsim.thresholds()
    rnd = random.Random(5)
    chunks = []
    for a in range(8):
        for j in rnd.sample(range(255), 12):
            key = thr[a][j]
            bits = int(f"{key:064b}"[::-1], 2)  # history with newest bit lowest
            chunks.append(bits.to_bytes(8, "big"))
    data = b"".join(chunks)
    stream = libs.oracle.compress_block(data)
    assert libs.hostsim.lane_decode_block(stream, len(data), 2, pair=pair) == data
