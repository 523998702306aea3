"""Second, independent CPU restatement of weath3rb0i's predict→update→code path.

TEST INFRASTRUCTURE ONLY.  Only tests/ may import this module (it is far too
slow for anything else: pure-Python loops, small cases only).  It exists
because the reference's Rust cannot be compiled in this image and the reference
holds no golden vectors for the model/history layers: two restatements written
separately (this one and oracle/w3_oracle.c) must agree byte for byte.

Parity status: coder + bit I/O pinned by the reference's KATs
(tests/test_oracle_kats.py replays them against this module too); model and
history layers "parity unpinned" by reference fixtures.

Citations are relative to /root/reference/.
"""
from __future__ import annotations

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF

# src/entropy_coding/arithmetic_coder.rs:4-9
Q1, Q2, Q3 = 0x40000000, 0x80000000, 0xC0000000

# src/models/ac_hash/stationary.rs:40-42
BOOK1 = (1, 50188, 62497, 15819, 22545, 31499, 22988, 29616)
# src/models/ac_hash/stationary.rs:44-46
ENWIK7 = (752, 50314, 58928, 21421, 24680, 30788, 24297, 32530)

MAGIC = b"w30i"  # src/main.rs:14


class SinkFull(Exception):
    """The io::Error EntropyWriter returns (src/history/ac_history.rs:64-66)."""


class ByteSink:
    """ACWriter — src/entropy_coding/io.rs:52-101."""

    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.fill = 0
        self.pending = 0

    def _push(self, b):  # io.rs:74-81
        self.acc = ((self.acc << 1) | b) & 0xFF
        self.fill = (self.fill + 1) % 8
        if self.fill == 0:
            self.out.append(self.acc)

    def write_bit(self, b):  # io.rs:70-89
        self._push(b)
        while self.pending:
            self.pending -= 1
            self._push(b ^ 1)

    def inc_parity(self):  # io.rs:66-68
        self.pending += 1

    def flush(self, state):  # io.rs:91-100
        self.write_bit((state >> 31) & 1)
        state = (state << 1) & M32
        while self.fill > 0:
            self.write_bit((state >> 31) & 1)
            state = (state << 1) & M32


class HashSink:
    """EntropyWriter — src/history/ac_history.rs:50-88."""

    def __init__(self, max_bits):
        self.state = 0
        self.count = 0
        self.pending = 0
        self.max_bits = max_bits

    def _push(self, b):  # ac_history.rs:63-71
        if self.count == self.max_bits:
            raise SinkFull
        self.state = (self.state >> 1) | (b << 31)
        self.count += 1

    def write_bit(self, b):  # ac_history.rs:58-80
        self._push(b)
        while self.pending:
            self.pending -= 1
            self._push(b ^ 1)

    def inc_parity(self):  # ac_history.rs:82-84
        self.pending += 1

    def value(self):  # ac_history.rs:45 (shift by 32 when count == 0 → 0)
        return 0 if self.count == 0 else self.state >> (32 - self.count)


class BitSource:
    """ACReader — src/entropy_coding/io.rs:7-49."""

    def __init__(self, data):
        self.data = bytes(data)
        self.at = 0
        self.cur = 0
        self.mask = 0

    def _byte(self):  # io.rs:18-27
        if self.at < len(self.data):
            v = self.data[self.at]
            self.at += 1
            return v
        return 0

    def read_bit(self):  # io.rs:31-38
        self.mask >>= 1
        if self.mask == 0:
            self.cur = self._byte()
            self.mask = 0x80
        return 1 if self.cur & self.mask else 0

    def read_u32(self):  # io.rs:40-48
        return int.from_bytes(bytes(self._byte() for _ in range(4)), "big")


def _split(lo, hi, prob):
    """lerp — src/entropy_coding/arithmetic_coder.rs:109-119."""
    scale = 1 if prob == 0 else prob << 16
    return lo + (((hi - lo) * scale) >> 32)


class Coder:
    """ArithmeticCoder — src/entropy_coding/arithmetic_coder.rs:36-106."""

    def __init__(self):
        self.lo, self.hi, self.x = 0, M32, 0

    def encode(self, bit, prob, sink):  # :41-65
        mid = _split(self.lo, self.hi, prob)
        if bit:
            self.hi = mid
        else:
            self.lo = mid + 1
        while (self.lo ^ self.hi) < Q2:  # top bits equal
            sink.write_bit(self.lo >> 31)
            self.lo = (self.lo << 1) & M32
            self.hi = ((self.hi << 1) & M32) | 1
        while self.lo >= Q1 and self.hi < Q3:
            sink.inc_parity()
            self.lo = (self.lo << 1) & 0x7FFFFFFF
            self.hi = ((self.hi << 1) & M32) | 0x80000001

    def start_decode(self, src):  # :75-78
        self.x = src.read_u32()

    def decode(self, prob, src):  # :80-105
        mid = _split(self.lo, self.hi, prob)
        bit = 1 if self.x <= mid else 0
        if bit:
            self.hi = mid
        else:
            self.lo = mid + 1
        while (self.lo ^ self.hi) < Q2:
            self.lo = (self.lo << 1) & M32
            self.hi = ((self.hi << 1) & M32) | 1
            self.x = ((self.x << 1) & M32) | src.read_bit()
        while self.lo >= Q1 and self.hi < Q3:
            self.lo = (self.lo << 1) & 0x7FFFFFFF
            self.hi = ((self.hi << 1) & M32) | 0x80000001
            self.x = (((self.x << 1) & M32) ^ Q2) | src.read_bit()
        return bit


def achash(bits, pos, table=BOOK1, max_bits=8):
    """ACHistory::hash — src/history/ac_history.rs:28-46."""
    coder = Coder()
    sink = HashSink(max_bits)
    a = pos & 7  # StationaryModel::align, stationary.rs:50-52
    try:
        for i in range(64):
            a = (a + 7) & 7  # stationary.rs:55
            coder.encode((bits >> i) & 1, table[a], sink)
    except SinkFull:
        pass
    return sink.value()


def counter_p(c0, c1):
    """Counter::p — src/models/counter.rs:13-18."""
    q = ((1 << 17) * (c1 + 1)) // (c0 + c1 + 2)
    return (q >> 1) + (q & 1)


def counter_update(c, bit):
    """Counter::update — src/models/counter.rs:20-26 (c is a 2-list)."""
    c[bit] += 1
    if c[bit] == 0xFFFF:
        c[0] = (c[0] >> 1) + (c[0] & 1)
        c[1] = (c[1] >> 1) + (c[1] & 1)


class OrderNEntropy:
    """OrderNEntropy<ACHistory<StationaryModel>> as built by init_model
    (src/main.rs:151; src/models/ordern_entropy.rs:5-45)."""

    def __init__(self, bits_in_context=11, alignment_bits=3, max_bits=8, table=BOOK1):
        self.stats = [[0, 0] for _ in range(1 << bits_in_context)]
        self.ctx = 0
        self.alignment = 0
        self.nctx = bits_in_context
        self.nalign = alignment_bits
        self.max_bits = max_bits
        self.table = table
        self.h_bits = 0
        self.h_pos = 0

    def predict(self):  # ordern_entropy.rs:28-30
        c = self.stats[self.ctx]
        return counter_p(c[0], c[1])

    def update(self, bit):  # models/mod.rs:28-31 → adapt then update
        counter_update(self.stats[self.ctx], bit)  # ordern_entropy.rs:32-34
        self.h_bits = ((self.h_bits << 1) | bit) & M64  # ac_history.rs:24
        self.h_pos += 1  # ac_history.rs:25
        self.alignment = (self.alignment + 1) & ((1 << self.nalign) - 1)
        h = achash(self.h_bits, self.h_pos, self.table, self.max_bits)
        h &= (1 << (self.nctx - self.nalign)) - 1
        self.ctx = (h << self.nalign) | self.alignment


def stationary_new(buf: bytes):
    """StationaryModel::new — src/models/ac_hash/stationary.rs:14-34."""
    model = [[0, 0] for _ in range(8)]
    for byte in buf:
        i = 7
        for k in range(7, -1, -1):
            i = (i + 1) & 7
            counter_update(model[i], (byte >> k) & 1)
    return tuple(counter_p(c[0], c[1]) for c in model)


class StatsSink:
    """ACStats — src/helpers.rs:60-90: counts bits, writes nothing; flush is a no-op."""

    def __init__(self):
        self.bit_count = 0
        self.rev_bits = 0

    def write_bit(self, b):  # helpers.rs:81-85
        self.bit_count += 1 + self.rev_bits
        self.rev_bits = 0

    def inc_parity(self):  # helpers.rs:77-79
        self.rev_bits += 1


def sweep_exec(buf: bytes, ctx_bits: int, alignment_bits: int, table=BOOK1) -> int:
    """exec() of the entropy-hashing-ac sweep (src/bin/entropy-hashing-ac/main.rs:24, 48-75):
    bits counted by ACStats; the printed csize is that // 8 (helpers.rs:71-73)."""
    sink, coder = StatsSink(), Coder()
    model = OrderNEntropy(ctx_bits, alignment_bits, ctx_bits - alignment_bits, table)
    for byte in buf:
        for k in range(7, -1, -1):
            bit = (byte >> k) & 1
            p = model.predict()
            model.update(bit)
            coder.encode(bit, p, sink)
    return sink.bit_count


def sweep_exec_ordern(buf: bytes, ctx_bits: int, alignment_bits: int) -> int:
    """compress() of the ordern sweep binary (src/bin/ordern/main.rs:66-81) with OrderN
    (src/models/ordern.rs:4-46): bits counted by ACStats."""
    stats = [[0, 0] for _ in range(1 << ctx_bits)]
    ctx = history = alignment = 0
    mask = (1 << (ctx_bits - alignment_bits)) - 1
    sink, coder = StatsSink(), Coder()
    for byte in buf:
        for k in range(7, -1, -1):
            bit = (byte >> k) & 1
            p = counter_p(*stats[ctx])
            counter_update(stats[ctx], bit)
            history = ((history << 1) | bit) & mask
            alignment = (alignment + 1) & ((1 << alignment_bits) - 1)
            ctx = (history << alignment_bits) | alignment
            coder.encode(bit, p, sink)
    return sink.bit_count


def compress_block(data: bytes, table=BOOK1, geo=(11, 3, 8)) -> bytes:
    """Stream after the 12-byte header — src/main.rs:99-111; geo = (bits_in_context,
    alignment_bits, max_bits) of OrderNEntropy::new / ACHistory::new (init_model: 11, 3, 8)."""
    sink, coder, model = ByteSink(), Coder(), OrderNEntropy(*geo, table=table)
    for byte in data:
        for k in range(7, -1, -1):
            bit = (byte >> k) & 1
            p = model.predict()
            model.update(bit)
            coder.encode(bit, p, sink)
    sink.flush(coder.hi)  # arithmetic_coder.rs:67-71
    return bytes(sink.out)


def decompress_block(stream: bytes, n_out: int, table=BOOK1, geo=(11, 3, 8)) -> bytes:
    """src/main.rs:127-140."""
    src, coder, model = BitSource(stream), Coder(), OrderNEntropy(*geo, table=table)
    coder.start_decode(src)
    out = bytearray()
    for _ in range(n_out):
        byte = 0
        for _ in range(8):
            p = model.predict()
            bit = coder.decode(p, src)
            model.update(bit)
            byte = (byte << 1) | bit
        out.append(byte)
    return bytes(out)


def compress_file(data: bytes) -> bytes:
    """src/main.rs:89-113 — magic + u64 BE length + stream."""
    return MAGIC + len(data).to_bytes(8, "big") + compress_block(data)


def decompress_file(blob: bytes) -> bytes:
    """src/main.rs:115-144."""
    if blob[:4] != MAGIC:
        raise ValueError("Magic numbers don't match up")  # main.rs:124
    n = int.from_bytes(blob[4:12], "big")
    return decompress_block(blob[12:], n)


def ac_encode_probs(data: bytes, probs) -> bytes:
    """compress() helper of arithmetic_coder_tests.rs:8-23."""
    src, sink, coder = BitSource(data), ByteSink(), Coder()
    for p in probs:
        coder.encode(src.read_bit(), p, sink)
    sink.flush(coder.hi)
    return bytes(sink.out)


def ac_decode_probs(stream: bytes, probs) -> bytes:
    """decompress() helper of arithmetic_coder_tests.rs:25-42 (the trailing
    0x00 the helper pops is never produced here)."""
    src, coder = BitSource(stream), Coder()
    coder.start_decode(src)
    out = ByteSink()
    for p in probs:
        out.write_bit(coder.decode(p, src))
    return bytes(out.out)


def ctx_trace(data: bytes):
    model = OrderNEntropy()
    tr = []
    for byte in data:
        for k in range(7, -1, -1):
            tr.append(model.ctx)
            model.update((byte >> k) & 1)
    return tr, model.ctx
