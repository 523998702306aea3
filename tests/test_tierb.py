"""the reference's state table (src/state_table/) and hash-slot
cells (src/hashmap.rs).  Both are dead code there, but the reference ships dumps of them —
docs/state_table/state_table.csv (3963 rows), aux_state_table.csv (990 rows) and the 60-row
index table of docs/hashslots.md:69-133 — committed here as tests/golden/tier_b_golden.json
(tests/golden/make_tierb_golden.py).  These rows are pinned by reference-issued data.

CPU part: the oracle and the product's statements (csrc/w3_tierb.cuh through tests/hostsim)
against the golden tables and against each other.  tests/test_gpu_tierb.py runs the device."""
import ctypes as C
import hashlib
import json
import os
import random

import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tier_b_golden.json")


@pytest.fixture(scope="module")
def gold():
    return json.load(open(GOLD))


def test_fixture_is_the_reference_dump(gold):
    """Where the reference tree is present (the build container) the fixture must equal its files."""
    ref = "/root/reference/docs"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this machine")
    for rel, want in gold["sha256"].items():
        assert hashlib.sha256(open(os.path.join(ref, rel), "rb").read()).hexdigest() == want, rel
    import csv

    for name, f in (("state_table", "state_table/state_table.csv"), ("aux_state_table", "state_table/aux_state_table.csv")):
        rows = list(csv.DictReader(open(os.path.join(ref, f))))
        assert [int(r["tr0"]) for r in rows] == gold[name]["tr0"]
        assert [int(r["tr1"]) for r in rows] == gold[name]["tr1"]
        assert [int(r["prob"]) for r in rows] == gold[name]["prob"]


def test_state_table_oracle_and_hostsim_equal_reference_csv(libs, gold):
    o, h = libs.oracle.lib, libs.hostsim.lib
    st = gold["state_table"]
    assert o.w3o_state_table_size() == len(st["prob"]) == 3963
    out = (C.c_uint16 * 3)()
    for s in range(3963):
        want = (st["tr0"][s], st["tr1"][s], st["prob"][s])
        assert (o.w3o_state_next(s, 0), o.w3o_state_next(s, 1), o.w3o_state_p(s)) == want, s
        h.hs_state_entry(0, s, out)
        assert tuple(out) == want, s
    at = gold["aux_state_table"]
    assert o.w3o_aux_state_table_size() == len(at["prob"]) == 990
    for s in range(990):
        want = (at["tr0"][s], at["tr1"][s], at["prob"][s])
        assert (o.w3o_aux_state_next(s, 0), o.w3o_aux_state_next(s, 1), o.w3o_aux_state_p(s)) == want, s
        h.hs_state_entry(1, s, out)
        assert tuple(out) == want, s


def test_state_table_structure(libs):
    """What naive.rs:17-48 promises: every transition stays inside the table, a 1 leads to a
    sub-table of odd index (b or d), a 0 to an even one, and p follows the counts."""
    o = libs.oracle.lib
    for s in range(3963):
        for bit in (0, 1):
            n = o.w3o_state_next(s, bit)
            assert n < 3963
            if s >= 3:
                assert ((n - 3) // 990) & 1 == bit
    states, nib = (C.c_uint16 * 4)(0, 1, 2, 3), 0b1010
    nxt, p4 = (C.c_uint16 * 4)(), (C.c_uint16 * 4)()
    o.w3o_state_next4(states, nib, nxt)  # mod.rs:5-12: bits of the nibble, MSB first
    assert list(nxt) == [o.w3o_state_next(0, 1), o.w3o_state_next(1, 0), o.w3o_state_next(2, 1), o.w3o_state_next(3, 0)]
    o.w3o_state_p4(states, p4)
    assert list(p4) == [o.w3o_state_p(s) for s in range(4)]


def test_slot_index_equals_reference_doc_table(libs, gold):
    o, h = libs.oracle.lib, libs.hostsim.lib
    assert len(gold["hashslots"]) == 60
    for sid, bit_id, nib_ctx, rel, ab, par in gold["hashslots"]:
        a, p = C.c_uint32(), C.c_int()
        o.w3o_slot_get_idx(sid, bit_id, nib_ctx, C.byref(a), C.byref(p))
        assert (a.value, "mask" if p.value else "shift") == (ab, par)
        idx = h.hs_slot_idx(sid, bit_id, nib_ctx)
        assert (idx >> 1, "mask" if idx & 1 else "shift") == (ab, par)
        assert rel == (1 << bit_id) - 1 + nib_ctx  # docs/hashslots.md:62
    # all 60 states of a cell are distinct 12-bit fields covering the 90 slot bytes exactly
    seen = set()
    for sid in range(4):
        for bit_id in range(4):
            for ctx in range(1 << bit_id):
                seen.add(h.hs_slot_idx(sid, bit_id, ctx))
    assert sorted(seen) == [3 * k for k in range(60)]


def test_cell_ops_hostsim_equal_oracle(libs):
    """Random set/get sequences on a 96-byte cell: same bytes, same states (hashmap.rs:86-128)."""
    o, h = libs.oracle.lib, libs.hostsim.lib
    rnd = random.Random(7)
    a, b = bytearray(96), bytearray(96)
    for i in range(96):
        a[i] = b[i] = rnd.getrandbits(8)
    ca, cb = (C.c_uint8 * 96).from_buffer(a), (C.c_uint8 * 96).from_buffer(b)
    for _ in range(5000):
        sid, bit_id = rnd.randrange(4), rnd.randrange(4)
        ctx = rnd.randrange(1 << bit_id)
        op = rnd.randrange(4)
        if op == 0:
            st = rnd.randrange(4096)
            o.w3o_slot_set_state(ca, sid, bit_id, ctx, st)
            h.hs_slot_set_state(cb, sid, bit_id, ctx, st)
            assert h.hs_slot_get_state(cb, sid, bit_id, ctx) == st
        elif op == 1:
            assert o.w3o_slot_get_state(ca, sid, bit_id, ctx) == h.hs_slot_get_state(cb, sid, bit_id, ctx)
        elif op == 2:
            nib = rnd.randrange(16)
            s = (C.c_uint16 * 4)(*(rnd.randrange(3963) for _ in range(4)))
            o.w3o_slot_set_nib(ca, sid, nib, s)
            h.hs_slot_set_nib(cb, sid, nib, s)
            g1, g2 = (C.c_uint16 * 4)(), (C.c_uint16 * 4)()
            o.w3o_slot_get_nib(ca, sid, nib, g1)
            h.hs_slot_get_nib(cb, sid, nib, g2)
            assert list(g1) == list(g2) == list(s)
        else:
            hsh = rnd.getrandbits(64)
            if rnd.random() < 0.5:  # make it hit one of the four check hashes
                k = rnd.randrange(4)
                hc = int.from_bytes(a[:6], "big")
                hsh = (hsh & ~0xFFF) | ((hc >> (12 * (3 - k))) & 0xFFF)
            assert o.w3o_cell_get_slot(ca, hsh) == h.hs_cell_get_slot(cb, hsh)
        assert a == b


def test_cell_get_slot_priority_and_miss(libs):
    """hashmap.rs:56-68: slot 3 is in the low 12 bits and is tested first; a miss is slot 1."""
    o = libs.oracle.lib
    cell = bytearray(96)
    cell[:6] = (0x111222333444).to_bytes(6, "big")
    c = (C.c_uint8 * 96).from_buffer(cell)
    assert [o.w3o_cell_get_slot(c, h) for h in (0x444, 0x333, 0x222, 0x111, 0x555)] == [3, 2, 1, 0, 1]
    cell[:6] = (0x777777777777).to_bytes(6, "big")
    assert o.w3o_cell_get_slot(c, 0xABC777) == 3
    assert bytes(cell[:6]) == (0x777777777777).to_bytes(6, "big")  # the miss path never stores the check hash


def test_hashmap_sizing_and_cell_index(libs):
    o, h = libs.oracle.lib, libs.hostsim.lib
    # HashMap::new (hashmap.rs:7-10): log_cell_count = (log2(size) - log2(96)) as u32
    for size, want in ((1 << 20, 13), (1 << 24, 17), (96 << 10, 10), ((96 << 10) - 1, 9), (1 << 30, 23)):
        assert o.w3o_hashmap_log_cell_count(size) == want, size
    rnd = random.Random(3)
    for _ in range(1000):
        hsh, lg = rnd.getrandbits(64), rnd.randrange(1, 40)
        assert o.w3o_hashmap_cell_index(hsh, lg) == h.hs_cell_index(hsh, lg) == hsh >> (64 - lg)


def test_opinion_mixer2(libs):
    """OpinionMixer2::mix (src/mixers/opinion_mixer2.rs:5-10): ties go to p1."""
    o, h = libs.oracle.lib, libs.hostsim.lib
    rnd = random.Random(1)
    for _ in range(5000):
        p1, p2 = rnd.randrange(65536), rnd.randrange(65536)
        want = p1 if abs(p1 - 32768) >= abs(p2 - 32768) else p2
        assert o.w3o_opinion_mix2(p1, p2) == h.hs_opinion_mix2(p1, p2) == want
