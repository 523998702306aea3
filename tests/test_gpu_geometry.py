"""SURVEY.md §8(f) ranks 2 and 3 on the GPU: the model geometry as a run-time parameter
(OrderNEntropy::new(bits_in_context, alignment_bits, ACHistory::new(max_bits, ..)) —
src/models/ordern_entropy.rs:15-24, src/history/ac_history.rs:16-18) with the counter table in HBM,
container version 3, and the parameter sweep of src/bin/entropy-hashing-ac/main.rs into the
ACStats bit counter (src/helpers.rs:60-90).  Bit-exact against the oracle."""
import os

import pytest

from corpus import w3_corpus
from weath3rb0i_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context([0])
    yield c
    c.close()


def _split_v3(blob):
    assert blob[:4] == b"w3Bk" and blob[4] == 3
    geo = (blob[5], blob[6], blob[7])
    n = int.from_bytes(blob[8:16], "big")
    bs = int.from_bytes(blob[16:20], "big")
    nb = int.from_bytes(blob[20:24], "big")
    table = tuple(int.from_bytes(blob[24 + 2 * i: 26 + 2 * i], "big") for i in range(8))
    lens = [int.from_bytes(blob[40 + 4 * i: 44 + 4 * i], "big") for i in range(nb)]
    at, streams = 40 + 4 * nb, []
    for ln in lens:
        streams.append(blob[at: at + ln])
        at += ln
    assert at == len(blob)
    return geo, n, bs, table, streams


@pytest.mark.parametrize("geo", [(16, 3, 13), (20, 3, 17), (12, 4, 8), (14, 2, 20), (8, 0, 8), (11, 3, 6), (22, 3, 19)])
def test_geometry_streams_equal_oracle(ctx, libs, geo):
    data = w3_corpus.generate(8192 * 3 + 1234, offset=1717)
    ctx.set_geometry(*geo)
    try:
        assert ctx.get_geometry() == geo
        blob = ctx.compress(data, 8192)
        g, n, bs, table, streams = _split_v3(blob)
        assert (g, n, bs, len(streams)) == (geo, len(data), 8192, 4)
        for i, s in enumerate(streams):
            assert s == libs.oracle.compress_block_geo(data[i * 8192: (i + 1) * 8192], *geo), (geo, i)
        with pytest.raises(capi.W3Error):  # the reference's file format cannot carry a geometry
            ctx.compress(data[:100], legacy=True)
    finally:
        ctx.set_geometry(11, 3, 8)
    assert ctx.decompress(blob) == data  # the geometry comes from the container, not from the context
    assert ctx.compress(data, 8192)[4] == 1  # (11, 3, 8) is init_model's: container version 1 again


def test_geometry_with_another_model_and_hostile_input(ctx, libs):
    import random

    from tests._libs import BOOK1, ENWIK7

    rnd = random.Random(2)
    data = bytes(rnd.getrandbits(8) for _ in range(5000)) + bytes(5000) + b"\xff" * 3000
    ctx.set_model(ENWIK7)
    ctx.set_geometry(18, 3, 15)
    try:
        blob = ctx.compress(data, 4096)
        geo, _, _, table, streams = _split_v3(blob)
        assert table == tuple(ENWIK7)
        for i, s in enumerate(streams):
            assert s == libs.oracle.compress_block_geo(data[i * 4096: (i + 1) * 4096], 18, 3, 15, ENWIK7), i
    finally:
        ctx.set_geometry(11, 3, 8)
        ctx.set_model(BOOK1)
    assert ctx.decompress(blob) == data


def test_sweep_equals_oracle(ctx, libs):
    """Every point of the reference's grid shape (ctx x align 0..4) on a small input, all at once."""
    data = w3_corpus.generate(6000, offset=99)
    points = [(c, a) for c in range(8, 17) for a in range(0, 5)] + [(20, 3), (24, 3), (30, 4)]
    bits, table = ctx.sweep(data, points)
    assert table == libs.oracle.stationary_new(data)  # StationaryModel::new(&buf), main.rs:19 of the sweep binary
    for (c, a), got in zip(points, bits):
        assert got == libs.oracle.sweep_exec(data, c, a, table), (c, a)
    # a caller-supplied table (StationaryModel::from_table)
    bits2, t2 = ctx.sweep(data, [(11, 3), (12, 3)], table=(1, 50188, 62497, 15819, 22545, 31499, 22988, 29616))
    assert bits2[0] == libs.oracle.sweep_exec(data, 11, 3, None)


def test_ordern_sweep_equals_oracle(ctx, libs):
    """The `ordern` binary's grid (src/bin/ordern/main.rs:16-33): raw bit history as context."""
    data = w3_corpus.generate(9000, offset=41)
    points = [(c, a) for c in range(8, 18) for a in range(0, 5)] + [(24, 3), (30, 0)]
    bits, _ = ctx.sweep(data, points, history="raw")
    for (c, a), got in zip(points, bits):
        assert got == libs.oracle.sweep_exec_ordern(data, c, a), (c, a)
    path = os.environ.get("W3_BOOK1")
    if path and os.path.exists(path):  # src/bin/ordern/book1.log:4,22
        b, _ = ctx.sweep(open(path, "rb").read(), [(8, 3), (11, 3)], history="raw")
        assert [v // 8 for v in b] == [444793, 388051]


def test_sweep_larger_input_sampled_points(ctx, libs):
    data = w3_corpus.generate(1 << 18, offset=5)
    points = [(c, a) for c in range(8, 23) for a in range(0, 5)]
    bits, table = ctx.sweep(data, points)
    for c, a in ((8, 0), (11, 3), (16, 2), (20, 3), (22, 4)):
        assert bits[points.index((c, a))] == libs.oracle.sweep_exec(data, c, a, table), (c, a)
    # more context helps on text, as in the reference's log (book1.log:22 vs :76)
    assert bits[points.index((20, 3))] < bits[points.index((11, 3))]


def test_book1_reference_csize_on_the_gpu_if_available(ctx):
    """src/bin/entropy-hashing-ac/book1.log:22,76 — the reference's own published csize for this stack;
    runs wherever $W3_BOOK1 points at Calgary book1 (768 771 bytes), skipped otherwise."""
    path = os.environ.get("W3_BOOK1")
    if not path or not os.path.exists(path):
        pytest.skip("set W3_BOOK1=/path/to/calgary/book1")
    data = open(path, "rb").read()
    bits, _ = ctx.sweep(data, [(11, 3), (20, 3)])
    assert [b // 8 for b in bits] == [353992, 262871]
