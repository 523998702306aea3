"""Parity tests proper: every stage and the whole path, through the C ABI on a
B200, against the CPU oracle — bit-exact."""
import hashlib
import random

import pytest

from corpus import w3_corpus
from tests._libs import BOOK1
from tests.test_oracle_kats import CODER_KATS
from weath3rb0i_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context([0])
    yield c
    c.close()


@pytest.fixture(params=["team", "many", "warp", "pair"], autouse=True)
def kernel_mode(request, ctx):
    """Every test runs against every kernel family (W3_OPT_KERNELS): `team` (one CTA per
    block: coder warp + model / speculator warps), `many` (the many-block regime: teams
    sharing a CTA for encode; for decode the split form — a coder warp and eight model warps
    per CTA of up to 8 blocks — as the launcher picks it for these block counts), `pair` (the
    same with the decoder's pair form, W3_OPT_LANE_FORM 2: two lanes per block in one warp,
    what the launcher picks beyond 8 blocks per SM) and `warp` (round 1's one warp per block)."""
    ctx.set_kernels("many" if request.param == "pair" else request.param)
    ctx.set_option(capi.W3_OPT_LANE_FORM, 2 if request.param == "pair" else 0)
    yield request.param
    ctx.set_option(capi.W3_OPT_LANE_FORM, 0)
    ctx.set_kernels("auto")


def split_container(blob):
    assert blob[:4] == b"w3Bk" and blob[4] == 1
    n = int.from_bytes(blob[8:16], "big")
    bs = int.from_bytes(blob[16:20], "big")
    nb = int.from_bytes(blob[20:24], "big")
    lens = [int.from_bytes(blob[24 + 4 * i : 28 + 4 * i], "big") for i in range(nb)]
    at, streams = 24 + 4 * nb, []
    for ln in lens:
        streams.append(blob[at : at + ln])
        at += ln
    assert at == len(blob)
    return n, bs, streams


# ---- stages -----------------------------------------------------------------


def test_achash_device_equals_oracle(ctx, libs):
    rnd = random.Random(21)
    bits, pos = [], []
    for i in range(200000):
        b = rnd.getrandbits(64)
        if i % 3 == 0:
            b &= (1 << rnd.randrange(1, 64)) - 1
        bits.append(b)
        pos.append(rnd.randrange(1 << 24))
    # keys at / around every threshold, incl. the singular ones (exact fallback)
    thr, _ = libs.hostsim.thresholds()
    for a in range(8):
        for j in range(255):
            for key in (thr[a][j] - 1, thr[a][j], thr[a][j] + 1):
                if 0 <= key < 1 << 64:
                    bits.append(int(f"{key:064b}"[::-1], 2))
                    pos.append(a)
    got = ctx.achash(bits, pos)
    for b, p, g in zip(bits, pos, got):
        assert g == libs.oracle.achash(b, p)[0], (hex(b), p)


def test_golden_hashes_device(ctx):
    import json, os

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tier_a_vectors.json")))
    bits = [int(b, 16) for b, _, _ in g["achash"]]
    pos = [p for _, p, _ in g["achash"]]
    assert ctx.achash(bits, pos) == [w for _, _, w in g["achash"]]


def test_sector_gather_probe_is_plausible(ctx):
    """bench.py's gather-roofline denominator: between a tenth of and the hard ceiling of
    HBM bandwidth / 32 B (6498.7e9 / 32 = 203 G sectors/s)."""
    rate = ctx.sector_gather_rate(1 << 30)
    assert 5e9 < rate < 260e9, rate


def test_counter_p_exhaustive_on_device(ctx):
    assert ctx.counter_p_mismatches() == 0


@pytest.mark.parametrize("name,data,probs,length", CODER_KATS, ids=[k[0] for k in CODER_KATS])
def test_reference_coder_kats_on_device(ctx, libs, name, data, probs, length):
    """src/entropy_coding/arithmetic_coder_tests.rs:44-161 through the device coder."""
    comp = ctx.ac_encode_probs(data, probs)
    assert len(comp) == length
    assert comp == libs.oracle.ac_encode_probs(data, probs)
    assert ctx.ac_decode_probs(comp, probs) == data


def test_device_coder_random_probs(ctx, libs):
    rnd = random.Random(4)
    for n in (1, 7, 8, 9, 1000, 20000):
        data = bytes(rnd.getrandbits(8) for _ in range((n + 7) // 8))
        probs = [rnd.choice([0, 1, 65535, 32768, rnd.randrange(65536)]) for _ in range(n)]
        comp = ctx.ac_encode_probs(data, probs)
        assert comp == libs.oracle.ac_encode_probs(data, probs)
        back = ctx.ac_decode_probs(comp, probs)
        assert back == libs.oracle.ac_decode_probs(comp, probs)


# ---- the whole path ----------------------------------------------------------


def test_golden_streams_legacy_file(ctx, libs):
    import json, os

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tier_a_vectors.json")))
    for v in g["streams"]:
        data = bytes.fromhex(v["input_hex"])
        blob = ctx.compress(data, legacy=True)
        assert blob == b"w30i" + len(data).to_bytes(8, "big") + bytes.fromhex(v["stream_hex"]), v["name"]
        assert len(blob) == v["file_len"]
        assert ctx.decompress(blob) == data


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 127, 128, 129, 1000, 4097, 70001])
def test_single_block_equals_oracle(ctx, libs, n):
    data = w3_corpus.generate(n, offset=777)
    blob = ctx.compress(data, legacy=True)
    assert blob == libs.oracle.compress_file(data)
    assert ctx.decompress(blob) == data


@pytest.mark.parametrize("block,n", [(4096, 4096 * 5), (4096, 4096 * 5 + 1), (4096, 4096 * 37 + 4095),
                                     (65536, 65536 * 9 + 12345), (1024, 1024 * 300 + 7), (4, 41)])
def test_blocks_equal_oracle_ragged(ctx, libs, block, n):
    data = w3_corpus.generate(n, offset=99)
    blob = ctx.compress(data, block)
    nn, bs, streams = split_container(blob)
    assert (nn, bs) == (n, block) and len(streams) == (n + block - 1) // block
    for i, s in enumerate(streams):
        assert s == libs.oracle.compress_block(data[i * block : (i + 1) * block]), i
    assert ctx.decompress(blob) == data


def test_split_decoder_group_sizes_and_rounds(ctx, kernel_mode):
    """The split decoder (one coder warp + eight model warps per CTA, W3_OPT_LANE_FORM 4) at every
    group size: groups the block count does not fill (lanes shadowing real blocks), a ragged last
    block that keeps stepping past its end, and more groups than SMs (several rounds per CTA)."""
    if kernel_mode != "many":
        pytest.skip("one kernel family: the option pins the decoder")
    cases = [(4096, 4096 * 37 + 4095), (256, 256 * 1300 + 1), (1024, 1024 * 300 + 7)]
    try:
        for block, n in cases:
            data = w3_corpus.generate(n, offset=5)
            ctx.set_option(capi.W3_OPT_LANE_FORM, 0)
            blob = ctx.compress(data, block)
            ctx.set_option(capi.W3_OPT_LANE_FORM, 4)
            for group in (1, 2, 3, 5, 7, 8):
                ctx.set_option(capi.W3_OPT_LANE_WARPS, group)
                assert ctx.decompress(blob) == data, (block, n, group)
    finally:
        ctx.set_option(capi.W3_OPT_LANE_WARPS, 0)
        ctx.set_option(capi.W3_OPT_LANE_FORM, 0)


def test_empty_input(ctx, libs):
    assert ctx.compress(b"", legacy=True) == libs.oracle.compress_file(b"") == b"w30i" + bytes(8) + b"\xff"
    assert ctx.decompress(b"w30i" + bytes(8) + b"\xff") == b""
    blob = ctx.compress(b"", 4096)
    assert len(blob) == 24 and ctx.decompress(blob) == b""


def test_hostile_inputs(ctx, libs):
    rnd = random.Random(9)
    cases = [
        bytes(rnd.getrandbits(8) for _ in range(50000)),  # incompressible: output > input
        bytes(50000), b"\xff" * 50000, bytes(range(256)) * 100,
        b"\x80" * 9000 + b"\x01" * 9000,
        b"e" * 70000,  # drives a counter through the 65535 halving (counter.rs:22-25)
        bytes((i * 37) & 0xFF for i in range(20000)),
    ]
    for data in cases:
        blob = ctx.compress(data, legacy=True)
        assert blob == libs.oracle.compress_file(data)
        assert ctx.decompress(blob) == data
        blob = ctx.compress(data, 8192)
        _, _, streams = split_container(blob)
        for i, s in enumerate(streams):
            assert s == libs.oracle.compress_block(data[i * 8192 : (i + 1) * 8192])
        assert ctx.decompress(blob) == data


def test_overflowed_blocks_are_recoded_alone(ctx, libs):
    """A stream that does not fit its scratch region is measured, not stored, and only those blocks
    are coded again into regions of exactly their size.  With the first-pass region cut to half a
    block, text fits (ratio 0.45) and the random and all-ones... blocks do not."""
    import random as _r

    rnd = _r.Random(77)
    block = 8192
    parts = []
    for i in range(12):
        parts.append(bytes(rnd.getrandbits(8) for _ in range(block)) if i in (2, 3, 9) else w3_corpus.generate(block, offset=i * block))
    data = b"".join(parts) + w3_corpus.generate(1000)
    want = ctx.compress(data, block)
    ctx.set_option(capi.W3_OPT_SCRATCH_PERMILLE, 500)
    try:
        got = ctx.compress(data, block)
    finally:
        ctx.set_option(capi.W3_OPT_SCRATCH_PERMILLE, 0)
    assert got == want
    _, _, streams = split_container(got)
    for i in (1, 2, 9, 12):
        assert streams[i] == libs.oracle.compress_block(data[i * block : (i + 1) * block]), i
    assert ctx.decompress(got) == data


def test_truncated_legacy_stream_zero_pads(ctx, libs):
    """ACReader yields 0 bits past EOF (io.rs:23-25): a cut file still decodes
    `len` bytes, identically to the reference algorithm."""
    data = w3_corpus.generate(3000)
    blob = ctx.compress(data, legacy=True)
    cut = blob[: 12 + (len(blob) - 12) // 2]
    assert ctx.decompress(cut) == libs.oracle.decompress_block(cut[12:], len(data))


def test_bad_magic_and_bad_container(ctx):
    with pytest.raises(capi.MagicMismatch):  # src/main.rs:123-124
        ctx.decompress(b"w31i" + bytes(8) + b"\xff")
    good = ctx.compress(w3_corpus.generate(10000), 4096)
    with pytest.raises(capi.W3Error):
        ctx.decompress(good[:-1])
    with pytest.raises(capi.W3Error):
        ctx.compress(b"abc", 6)  # block size must be a multiple of 4


def test_config1_1mib_single_block(ctx, libs):
    """BASELINE.json configs[0]: reference `t` of the 1 MiB file, single block."""
    data = w3_corpus.generate(1 << 20)
    assert hashlib.sha256(data).hexdigest() == w3_corpus.SHA256_1MIB
    blob = ctx.compress(data, legacy=True)
    assert blob == libs.oracle.compress_file(data)
    assert ctx.decompress(blob) == data


def test_config2_64mib_every_block_identical(ctx, libs):
    """BASELINE.json configs[1]: 64 MiB, 1 MiB blocks, per-block byte-identical
    check.  The oracle side runs multi-threaded (w3o_compress_blocks_mt)."""
    import ctypes as C
    import os

    n, block = 64 << 20, 1 << 20
    buf = C.create_string_buffer(n)
    w3_corpus.generate_into(C.addressof(buf), n)
    data = buf.raw
    blob = ctx.compress(data, block)
    _, _, streams = split_container(blob)
    assert len(streams) == 64
    threads = min(64, len(os.sched_getaffinity(0)))
    if threads >= 16:
        stride = block + block // 8 + 64
        comp = C.create_string_buffer(stride * 64)
        lens = (C.c_uint32 * 64)()
        assert libs.oracle.lib.w3o_compress_blocks_mt(C.addressof(buf), n, block, C.addressof(comp), stride,
                                                      C.addressof(lens), threads) == 0
        for i in range(64):
            assert streams[i] == comp.raw[i * stride : i * stride + lens[i]], i
    else:  # few host cores: a deterministic sample of blocks
        for i in (0, 1, 31, 63):
            assert streams[i] == libs.oracle.compress_block(data[i * block : (i + 1) * block]), i
    assert ctx.decompress(blob) == data


def test_result_independent_of_block_grouping(ctx):
    """Blocks are independent: coding a sub-range alone gives the same streams
    (what sharding over GPUs relies on)."""
    data = w3_corpus.generate(4096 * 40)
    _, _, all_streams = split_container(ctx.compress(data, 4096))
    _, _, part = split_container(ctx.compress(data[4096 * 13 : 4096 * 29], 4096))
    assert all_streams[13:29] == part


def test_join_of_gpu_parts_equals_whole(ctx):
    """Host-side container join (weath3rb0i_b200/sharding.py) over GPU-coded parts."""
    from weath3rb0i_b200 import sharding

    data = w3_corpus.generate(4096 * 21 + 100)
    whole = ctx.compress(data, 4096)
    parts = []
    for r in range(3):
        lo, hi = sharding.byte_range(len(data), 4096, 3, r)
        parts.append(ctx.compress(data[lo:hi], 4096))
    assert sharding.join_containers(parts) == whole


def test_two_devices_in_one_context_give_the_same_container(ctx):
    """Blocks sharded over 2 GPUs by the library (one host thread per device)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    data = w3_corpus.generate(65536 * 7 + 333)
    c2 = capi.Context([0, 1])
    try:
        blob = c2.compress(data, 65536)
        assert blob == ctx.compress(data, 65536)
        assert c2.decompress(blob) == data
    finally:
        c2.close()


def test_256mib_round_trip_and_sampled_blocks(ctx, libs, kernel_mode):
    """Larger than one wave of CTAs (256 blocks on 148 SMs): lossless round trip
    of the whole container, byte identity with the oracle on a deterministic
    sample of blocks, and a container checksum that must not depend on the
    kernel family that produced it."""
    import ctypes as C

    if kernel_mode != "team":
        pytest.skip("run once, with the launcher's own choice of kernels")
    n, block = 256 << 20, 1 << 20
    buf = C.create_string_buffer(n)
    w3_corpus.generate_into(C.addressof(buf), n)
    data = buf.raw
    ctx.set_kernels("auto")  # teams sharing CTAs for encode, one lane per block for decode at this block count
    blob = ctx.compress(data, block)
    _, _, streams = split_container(blob)
    assert len(streams) == 256
    for i in (0, 97, 148, 255):
        assert streams[i] == libs.oracle.compress_block(data[i * block : (i + 1) * block]), i
    assert ctx.decompress(blob) == data
    for mode in ("team", "warp"):
        ctx.set_kernels(mode)
        assert hashlib.sha256(ctx.compress(data, block)).digest() == hashlib.sha256(blob).digest(), mode
