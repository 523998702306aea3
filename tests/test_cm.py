"""The hashed context-mixing stack "w3cm"  on the CPU.

PARITY UNPINNED by the reference: it has no logistic mixer, no APM, and no caller of its hash
map or state table, so there is nothing there to be identical to.  The specification is
oracle/w3_cm_oracle.c; these tests check that it is a lossless codec, that it uses the
reference's own pieces the way the reference defines them (cell layout, state table, coder),
and that the product's helper statements (csrc/w3_cm.cuh through tests/hostsim) agree with it.
tests/test_gpu_cm.py compares the kernels with this oracle per block."""
import ctypes as C
import random

import pytest

from corpus import w3_corpus


def test_squash_stretch(libs):
    o, h = libs.oracle.lib, libs.hostsim.lib
    prev = -1
    for d in range(-2200, 2201):
        v = o.w3o_cm_squash(d)
        assert v == h.hs_cm_squash(d) and 0 <= v <= 4095 and v >= prev
        prev = v
    assert o.w3o_cm_squash(0) == 2047 and o.w3o_cm_squash(-3000) == 0 and o.w3o_cm_squash(3000) == 4095
    for p in range(4096):
        s = o.w3o_cm_stretch(p)
        assert s == h.hs_cm_stretch(p) and -2047 <= s <= 2047
        if 1 <= p <= 4094:  # stretch is the inverse table: the smallest d whose squash reaches p
            assert o.w3o_cm_squash(s) >= p and (s == -2047 or o.w3o_cm_squash(s - 1) < p)


def test_hash_state_bits_and_training_step(libs):
    o, h = libs.oracle.lib, libs.hostsim.lib
    rnd = random.Random(8)
    for _ in range(5000):
        last8, m, low, hi = rnd.getrandbits(64), rnd.randrange(6), rnd.getrandbits(1), rnd.randrange(16)
        assert o.w3o_cm_hash(last8, m, low, hi) == h.hs_cm_hash(last8, m, low, hi)
    # the model's order is the number of bytes it sees
    assert o.w3o_cm_hash(0x1122334455667788, 0, 0, 0) == o.w3o_cm_hash(0xFFFFFFFFFFFFFF88, 0, 0, 0)
    assert o.w3o_cm_hash(0x1122334455667788, 1, 0, 0) != o.w3o_cm_hash(0xFFFFFFFFFFFFFF88, 1, 0, 0)
    for s in range(3963):
        assert o.w3o_cm_state_bits(s) == h.hs_cm_state_bits(s)
    assert [o.w3o_cm_state_bits(s) for s in (0, 1, 2, 3, 4)] == [0, 1, 1, 2, 3]
    for _ in range(2000):
        w, st, err = rnd.randrange(-(1 << 18), 1 << 18), rnd.randrange(-2047, 2048), rnd.randrange(-24570, 24571)
        want = max(-(1 << 18), min(1 << 18, w + ((st * err + (1 << 15)) >> 16)))
        assert h.hs_cm_train(w, st, err) == want


def test_find_slot_hit_order_and_replacement(libs):
    """Hits follow Cell::get_slot (src/hashmap.rs:56-63: slot 3 first); a miss clears the slot whose first
    state has seen the fewest bits and stores the check hash there — the reference leaves that open (:64-68)."""
    o, h = libs.oracle.lib, libs.hostsim.lib
    rnd = random.Random(4)
    for _ in range(3000):
        a = bytearray(rnd.getrandbits(8) for _ in range(96))
        for sid in range(4):  # plausible first states
            o.w3o_slot_set_state((C.c_uint8 * 96).from_buffer(a), sid, 0, 0, rnd.choice([0, 1, 2, 3, 50, 700, 3000]))
        b = bytearray(a)
        hsh = rnd.getrandbits(64)
        if rnd.random() < 0.5:
            hc = int.from_bytes(a[:6], "big")
            hsh = (hsh & ~0xFFF) | ((hc >> (12 * rnd.randrange(4))) & 0xFFF)
        ca, cb = (C.c_uint8 * 96).from_buffer(a), (C.c_uint8 * 96).from_buffer(b)
        assert o.w3o_cm_find_slot(ca, hsh) == h.hs_cm_find_slot(cb, hsh)
        assert a == b
    cell = bytearray(96)
    c = (C.c_uint8 * 96).from_buffer(cell)
    cell[:6] = (0x111222333444).to_bytes(6, "big")
    for sid, st in enumerate((700, 3, 50, 1)):  # slot 3's first state has seen the fewest bits
        o.w3o_slot_set_state(c, sid, 0, 0, st)
    assert o.w3o_cm_find_slot(c, 0xABC555) == 3
    assert int.from_bytes(cell[:6], "big") == 0x111222333555
    assert all(o.w3o_slot_get_state(c, 3, b, k) == 0 for b in range(4) for k in range(1 << b))
    assert o.w3o_slot_get_state(c, 2, 0, 0) == 50  # the neighbours keep their states (shared boundary byte 67)


@pytest.mark.parametrize("n", [0, 1, 2, 17, 1000, 30000])
def test_round_trip_and_it_beats_the_single_model_stack(libs, n):
    data = w3_corpus.generate(n, offset=55)
    for log_cells in (6, 12, 16):
        c = libs.oracle.cm_compress_block(data, log_cells)
        assert libs.oracle.cm_decompress_block(c, n, log_cells) == data
    if n >= 30000:  # six hashed orders + mixer + APM against init_model's one 11-bit table
        assert len(libs.oracle.cm_compress_block(data, 16)) < 0.95 * len(libs.oracle.compress_block(data))


def test_round_trip_hostile(libs):
    rnd = random.Random(6)
    for data in (bytes(rnd.getrandbits(8) for _ in range(20000)), bytes(40000), b"\xff" * 40000, bytes(range(256)) * 60,
                 b"ab" * 20000):
        c = libs.oracle.cm_compress_block(data, 8)  # 256 cells: constant replacement pressure
        assert libs.oracle.cm_decompress_block(c, len(data), 8) == data
