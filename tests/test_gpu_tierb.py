"""SURVEY.md §8 rows a17 / a18 on the device: csrc/w3_tierb.cuh through the C ABI against the
reference's own dumps (tests/golden/tier_b_golden.json) and against the oracle."""
import ctypes as C
import json
import os
import random

import pytest

from weath3rb0i_b200 import capi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "tier_b_golden.json")


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context([0])
    yield c
    c.close()


def test_device_state_table_equals_reference_csv(ctx):
    g = json.load(open(GOLD))
    for which, name in ((0, "state_table"), (1, "aux_state_table")):
        n0, n1, p = ctx.state_table(which)
        assert n0 == g[name]["tr0"] and n1 == g[name]["tr1"] and p == g[name]["prob"], name


def test_device_slot_index_equals_reference_doc_table(ctx):
    g = json.load(open(GOLD))
    cells = bytearray(96)
    res = ctx.hashslot_ops([(0, 0, sid, bit_id, ctx_) for sid, bit_id, ctx_, _, _, _ in g["hashslots"]], cells)
    for (abs_idx, parity), row in zip(res, g["hashslots"]):
        assert (abs_idx, "mask" if parity else "shift") == (row[4], row[5]), row


def test_device_cell_ops_equal_oracle(ctx, libs):
    """A random interleaving of set_state / get_state / set_nib / get_nib / get_slot / cell index over
    several cells, replayed in order on the device and by the oracle: same results, same bytes."""
    o = libs.oracle.lib
    rnd = random.Random(23)
    n_cells = 5
    dev = bytearray(rnd.getrandbits(8) for _ in range(96 * n_cells))
    host = bytearray(dev)
    views = [(C.c_uint8 * 96).from_buffer(host, 96 * i) for i in range(n_cells)]
    ops, want = [], []
    for _ in range(4000):
        cell, sid, bit_id = rnd.randrange(n_cells), rnd.randrange(4), rnd.randrange(4)
        nctx = rnd.randrange(1 << bit_id)
        code = rnd.choice([1, 1, 2, 2, 3, 4, 5, 6])
        if code == 1:
            ops.append((1, cell, sid, bit_id, nctx))
            want.append((o.w3o_slot_get_state(views[cell], sid, bit_id, nctx), 0))
        elif code == 2:
            st = rnd.randrange(4096)
            ops.append((2, cell, sid | (bit_id << 8) | (nctx << 16), st, 0))
            o.w3o_slot_set_state(views[cell], sid, bit_id, nctx, st)
            want.append((0, 0))
        elif code == 3:
            hsh = rnd.getrandbits(64)
            if rnd.random() < 0.6:
                hc = int.from_bytes(host[96 * cell : 96 * cell + 6], "big")
                hsh = (hsh & ~0xFFF) | ((hc >> (12 * rnd.randrange(4))) & 0xFFF)
            ops.append((3, cell, hsh & 0xFFFFFFFF, hsh >> 32, 0))
            want.append((o.w3o_cell_get_slot(views[cell], hsh), 0))
        elif code == 4:
            nib = rnd.randrange(16)
            s = [rnd.randrange(3963) for _ in range(4)]
            ops.append((4, cell, sid | (nib << 8), s[0] | (s[1] << 16), s[2] | (s[3] << 16)))
            o.w3o_slot_set_nib(views[cell], sid, nib, (C.c_uint16 * 4)(*s))
            want.append((0, 0))
        elif code == 5:
            nib = rnd.randrange(16)
            g = (C.c_uint16 * 4)()
            o.w3o_slot_get_nib(views[cell], sid, nib, g)
            ops.append((5, cell, sid | (nib << 8), 0, 0))
            want.append((g[0] | (g[1] << 16), g[2] | (g[3] << 16)))
        else:
            hsh, lg = rnd.getrandbits(64), rnd.randrange(1, 40)
            ix = o.w3o_hashmap_cell_index(hsh, lg)
            ops.append((6, cell, hsh & 0xFFFFFFFF, hsh >> 32, lg))
            want.append((ix & 0xFFFFFFFF, ix >> 32))
    del views
    got = ctx.hashslot_ops(ops, dev)
    assert got == want
    assert dev == host
