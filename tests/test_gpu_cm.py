"""The hashed context-mixing stack "w3cm" on the GPU: one warp per
block, lanes = context models, cells in HBM.  PARITY UNPINNED by the reference (no such model
there); the bar is identity with the specification, oracle/w3_cm_oracle.c, per block, and a
lossless round trip."""
import random

import pytest

from corpus import w3_corpus
from weath3rb0i_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context([0])
    yield c
    c.close()


def split_cm(blob):
    assert blob[:4] == b"w3Cm" and blob[4] == 1 and blob[5] == 6
    log_cells = blob[6]
    n = int.from_bytes(blob[8:16], "big")
    bs = int.from_bytes(blob[16:20], "big")
    nb = int.from_bytes(blob[20:24], "big")
    lens = [int.from_bytes(blob[24 + 4 * i: 28 + 4 * i], "big") for i in range(nb)]
    at, streams = 24 + 4 * nb, []
    for ln in lens:
        streams.append(blob[at: at + ln])
        at += ln
    assert at == len(blob)
    return log_cells, n, bs, streams


@pytest.mark.parametrize("log_cells", [4, 8, 12, 16])
@pytest.mark.parametrize("block,n", [(4096, 4096 * 3 + 17), (16384, 16384 * 2), (65536, 70001), (1024, 5)])
def test_streams_equal_oracle(ctx, libs, log_cells, block, n):
    """Small tables (16 and 256 cells) make several models of a nibble share a cell all the time:
    the lanes then take turns on one staging row, in the oracle's model order."""
    data = w3_corpus.generate(n, offset=321)
    blob = ctx.cm_compress(data, block, log_cells)
    lc, nn, bs, streams = split_cm(blob)
    assert (lc, nn, bs) == (log_cells, n, block) and len(streams) == (n + block - 1) // block
    for i, s in enumerate(streams):
        assert s == libs.oracle.cm_compress_block(data[i * block: (i + 1) * block], log_cells), (log_cells, i)
    assert ctx.cm_decompress(blob) == data


def test_empty_and_hostile(ctx, libs):
    assert ctx.cm_decompress(ctx.cm_compress(b"", 4096, 10)) == b""
    rnd = random.Random(12)
    cases = [bytes(rnd.getrandbits(8) for _ in range(30000)), bytes(30000), b"\xff" * 30000, bytes(range(256)) * 50,
             b"ab" * 9000 + b"\x00\xff" * 3000]
    for data in cases:
        blob = ctx.cm_compress(data, 8192, 9)
        _, _, _, streams = split_cm(blob)
        for i, s in enumerate(streams):
            assert s == libs.oracle.cm_compress_block(data[i * 8192: (i + 1) * 8192], 9)
        assert ctx.cm_decompress(blob) == data


def test_many_blocks_round_trip_and_ratio(ctx, libs):
    """64 blocks of 128 KiB (8 MiB): a sample of blocks against the oracle, round trip, and the ratio the
    stack buys over init_model's single table on the same blocks."""
    n, block = 8 << 20, 128 << 10
    data = w3_corpus.generate(n)
    blob = ctx.cm_compress(data, block, 16)
    _, _, _, streams = split_cm(blob)
    assert len(streams) == 64
    for i in (0, 31, 63):
        assert streams[i] == libs.oracle.cm_compress_block(data[i * block: (i + 1) * block], 16), i
    assert ctx.cm_decompress(blob) == data
    assert len(blob) < 0.9 * len(ctx.compress(data, block))


def test_bad_container(ctx):
    with pytest.raises(capi.MagicMismatch):
        ctx.cm_decompress(b"w3Bk" + bytes(30))
    good = ctx.cm_compress(w3_corpus.generate(5000), 4096, 8)
    with pytest.raises(capi.W3Error):
        ctx.cm_decompress(good[:-1])
