"""other StationaryModels than for_book1 —
StationaryModel::from_table (stationary.rs:36-46) and StationaryModel::new
(:14-34, computed on the GPU, Counter::update halvings included), the model
carried by container version 2.  Bit-exact against the oracle."""
import random

import pytest

from corpus import w3_corpus
from tests._libs import BOOK1, ENWIK7
from weath3rb0i_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context([0])
    yield c
    c.close()


@pytest.fixture(params=["team", "many", "warp"], autouse=True)
def kernel_mode(request, ctx):
    ctx.set_kernels(request.param)
    yield request.param
    ctx.set_kernels("auto")


def _split_v2(blob):
    assert blob[:4] == b"w3Bk" and blob[4] == 2
    n = int.from_bytes(blob[8:16], "big")
    bs = int.from_bytes(blob[16:20], "big")
    nb = int.from_bytes(blob[20:24], "big")
    table = tuple(int.from_bytes(blob[24 + 2 * i : 26 + 2 * i], "big") for i in range(8))
    lens = [int.from_bytes(blob[40 + 4 * i : 44 + 4 * i], "big") for i in range(nb)]
    at, streams = 40 + 4 * nb, []
    for ln in lens:
        streams.append(blob[at : at + ln])
        at += ln
    assert at == len(blob)
    return n, bs, table, streams


def test_train_model_equals_stationary_new(ctx, libs, kernel_mode):
    if kernel_mode != "team":
        pytest.skip("training does not depend on the codec kernel family")
    rnd = random.Random(31)
    cases = [
        b"", b"a", w3_corpus.generate(1000), w3_corpus.generate(65534), w3_corpus.generate(65535),
        w3_corpus.generate(65536), w3_corpus.generate(3 << 20, offset=99),  # every position halves, several times
        bytes(rnd.getrandbits(8) for _ in range(200001)),
        bytes(200000), b"\xff" * 200000, b"\x55" * 131071, b"\xaa" * 131073,
        bytes(65530) + b"\x80" * 70000 + bytes(1030),  # halving lands mid-chunk, at various offsets
    ]
    for data in cases:
        assert ctx.train_model(data) == libs.oracle.stationary_new(data), len(data)


def test_from_table_models_code_like_the_oracle(ctx, libs):
    data = w3_corpus.generate(4096 * 9 + 77, offset=31337)
    try:
        for table in (ENWIK7, libs.oracle.stationary_new(data), (1, 2, 65535, 32768, 40000, 3, 65534, 100)):
            ctx.set_model(table)
            assert ctx.get_model() == tuple(table)
            blob = ctx.compress(data, 4096)
            n, bs, t, streams = _split_v2(blob)
            assert (n, bs, t) == (len(data), 4096, tuple(table))
            for i, s in enumerate(streams):
                assert s == libs.oracle.compress_block(data[i * 4096 : (i + 1) * 4096], table), i
            # a fresh model in the context: the container brings its own
            ctx.set_model(BOOK1)
            assert ctx.decompress(blob) == data
            assert ctx.get_model() == tuple(table)
            with pytest.raises(capi.W3Error):  # the reference's file format implies for_book1
                ctx.compress(data[:100], legacy=True)
    finally:
        ctx.set_model(BOOK1)
    # version 1 / legacy streams switch the context back to for_book1
    ctx.set_model(ENWIK7)
    v1 = b"w30i" + (3).to_bytes(8, "big") + libs.oracle.compress_block(b"abc")
    assert ctx.decompress(v1) == b"abc" and ctx.get_model() == BOOK1


def test_achash_follows_the_model(ctx, libs, kernel_mode):
    if kernel_mode != "team":
        pytest.skip("hash self-test does not depend on the codec kernel family")
    rnd = random.Random(3)
    bits = [rnd.getrandbits(64) for _ in range(20000)]
    pos = [rnd.randrange(64) for _ in range(20000)]
    try:
        ctx.set_model(ENWIK7)
        got = ctx.achash(bits, pos)
        for b, p, g in zip(bits, pos, got):
            assert g == libs.oracle.achash(b, p, ENWIK7, 8)[0]
    finally:
        ctx.set_model(BOOK1)


def test_trained_model_compresses_its_own_text_better(ctx):
    """What the reference's sweep does (src/bin/entropy-hashing-ac/main.rs:19):
    train on the input, then code it."""
    data = w3_corpus.generate(1 << 20)
    base = len(ctx.compress(data, 1 << 18))
    try:
        ctx.set_model(ctx.train_model(data))
        trained = ctx.compress(data, 1 << 18)
        assert ctx.decompress(trained) == data
        assert len(trained) < base * 1.02
    finally:
        ctx.set_model(BOOK1)
