"""ctypes loaders for the test-side native libraries.

oracle/libw3oracle.so  — the CPU restatement of the reference (the checker)
tests/hostsim/…        — w3_core.cuh compiled for the CPU (what the kernels execute)
Both are built on demand with gcc/g++ (seconds).  The product library
(weath3rb0i_b200/libw3codec.so) is loaded separately by weath3rb0i_b200.capi.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOOK1 = (1, 50188, 62497, 15819, 22545, 31499, 22988, 29616)
ENWIK7 = (752, 50314, 58928, 21421, 24680, 30788, 24297, 32530)


def _newer(target, *srcs):
    return os.path.exists(target) and all(
        os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs
    )


def build_oracle():
    d = os.path.join(ROOT, "oracle")
    so = os.path.join(d, "libw3oracle.so")
    if not _newer(so, *(os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".c", ".h")) or f == "Makefile")):
        subprocess.check_call(["make", "-s", "-C", d])
    return so


def build_hostsim():
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libw3hostsim.so")
    csrc = os.path.join(ROOT, "weath3rb0i_b200", "csrc")
    src = os.path.join(d, "w3_hostsim.cpp")
    if not _newer(so, src, *(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh"))):
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-Wall", "-shared", "-fPIC", "-o", so, src]
        )
    return so


def load_oracle():
    o = C.CDLL(build_oracle())
    sz, u8p, u16p, u32p = C.c_size_t, C.c_char_p, C.POINTER(C.c_uint16), C.POINTER(C.c_uint32)
    o.w3o_counter_p.restype = C.c_uint16
    o.w3o_counter_p.argtypes = [C.c_uint16, C.c_uint16]
    o.w3o_counter_update.argtypes = [u16p, C.c_int]
    o.w3o_achash.restype = C.c_uint32
    o.w3o_achash.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_int)]
    o.w3o_achash_ex.restype = C.c_uint32
    o.w3o_achash_ex.argtypes = [C.c_uint64, C.c_uint64, u16p, C.c_int, C.POINTER(C.c_int)]
    o.w3o_writer_new.restype = C.c_void_p
    o.w3o_writer_new.argtypes = [C.c_void_p, sz]
    for f in ("w3o_writer_free", "w3o_writer_inc_parity"):
        getattr(o, f).argtypes = [C.c_void_p]
    o.w3o_writer_write_bit.argtypes = [C.c_void_p, C.c_int]
    o.w3o_writer_flush.argtypes = [C.c_void_p, C.c_uint32]
    o.w3o_writer_len.restype = sz
    o.w3o_writer_len.argtypes = [C.c_void_p]
    o.w3o_reader_new.restype = C.c_void_p
    o.w3o_reader_new.argtypes = [C.c_void_p, sz]
    o.w3o_reader_free.argtypes = [C.c_void_p]
    o.w3o_reader_read_bit.argtypes = [C.c_void_p]
    o.w3o_reader_read_u32.restype = C.c_uint32
    o.w3o_reader_read_u32.argtypes = [C.c_void_p]
    o.w3o_ac_encode_probs.argtypes = [u8p, sz, u16p, sz, C.c_void_p, sz, C.POINTER(sz)]
    o.w3o_ac_decode_probs.argtypes = [u8p, sz, u16p, sz, C.c_void_p, sz]
    o.w3o_compress_block.argtypes = [u8p, sz, C.c_void_p, sz, C.POINTER(sz)]
    o.w3o_decompress_block.argtypes = [u8p, sz, C.c_void_p, sz]
    o.w3o_compress_block_ex.argtypes = [u8p, sz, u16p, C.c_void_p, sz, C.POINTER(sz)]
    o.w3o_decompress_block_ex.argtypes = [u8p, sz, u16p, C.c_void_p, sz]
    o.w3o_stationary_new.argtypes = [u8p, sz, u16p]
    o.w3o_compress_block_geo.argtypes = [u8p, sz, u16p, C.c_int, C.c_int, C.c_int, C.c_void_p, sz, C.POINTER(sz)]
    o.w3o_decompress_block_geo.argtypes = [u8p, sz, u16p, C.c_int, C.c_int, C.c_int, C.c_void_p, sz]
    o.w3o_sweep_exec.argtypes = [u8p, sz, u16p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    o.w3o_sweep_exec_ordern.argtypes = [u8p, sz, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    o.w3o_compress_file.argtypes = [u8p, sz, C.c_void_p, sz, C.POINTER(sz)]
    o.w3o_decompress_file.argtypes = [u8p, sz, C.c_void_p, sz, C.POINTER(sz)]
    o.w3o_ctx_trace.argtypes = [u8p, sz, u16p]
    o.w3o_state_next.restype = o.w3o_state_p.restype = o.w3o_aux_state_next.restype = o.w3o_aux_state_p.restype = C.c_uint16
    o.w3o_state_next.argtypes = o.w3o_aux_state_next.argtypes = [C.c_uint16, C.c_int]
    o.w3o_state_p.argtypes = o.w3o_aux_state_p.argtypes = [C.c_uint16]
    o.w3o_state_next4.argtypes = [u16p, C.c_uint8, u16p]
    o.w3o_state_p4.argtypes = [u16p, u16p]
    o.w3o_hashmap_log_cell_count.restype = C.c_uint32
    o.w3o_hashmap_log_cell_count.argtypes = [C.c_uint64]
    o.w3o_hashmap_cell_index.restype = C.c_uint64
    o.w3o_hashmap_cell_index.argtypes = [C.c_uint64, C.c_uint32]
    o.w3o_cell_get_slot.argtypes = [C.c_void_p, C.c_uint64]
    o.w3o_slot_get_idx.argtypes = [C.c_int, C.c_int, C.c_int, u32p, C.POINTER(C.c_int)]
    o.w3o_slot_get_state.restype = C.c_uint16
    o.w3o_slot_get_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    o.w3o_slot_set_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint16]
    o.w3o_slot_set_nib.argtypes = [C.c_void_p, C.c_int, C.c_uint8, u16p]
    o.w3o_slot_get_nib.argtypes = [C.c_void_p, C.c_int, C.c_uint8, u16p]
    o.w3o_opinion_mix2.restype = C.c_uint16
    o.w3o_opinion_mix2.argtypes = [C.c_uint16, C.c_uint16]
    o.w3o_cm_squash.argtypes = o.w3o_cm_stretch.argtypes = [C.c_int]
    o.w3o_cm_state_bits.argtypes = [C.c_uint16]
    o.w3o_cm_hash.restype = C.c_uint64
    o.w3o_cm_hash.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_uint]
    o.w3o_cm_find_slot.argtypes = [C.c_void_p, C.c_uint64]
    o.w3o_cm_compress_block.argtypes = [u8p, sz, C.c_int, C.c_void_p, sz, C.POINTER(sz)]
    o.w3o_cm_decompress_block.argtypes = [u8p, sz, C.c_int, C.c_void_p, sz]
    o.w3o_compress_blocks_mt.argtypes = [C.c_void_p, sz, sz, C.c_void_p, sz, C.c_void_p, C.c_int]
    o.w3o_decompress_blocks_mt.argtypes = [C.c_void_p, sz, C.c_void_p, C.c_void_p, sz, sz, C.c_int]
    return o


class Oracle:
    """Pythonic face of libw3oracle.so (checker only)."""

    def __init__(self):
        self.lib = load_oracle()

    def compress_block(self, data: bytes, table=None) -> bytes:
        n = len(data)
        cap = 17 * n + 64
        out = C.create_string_buffer(cap)
        ol = C.c_size_t()
        t = (C.c_uint16 * 8)(*table) if table is not None else None
        rc = self.lib.w3o_compress_block_ex(data, n, t, out, cap, C.byref(ol))
        assert rc == 0, rc
        return out.raw[: ol.value]

    def decompress_block(self, stream: bytes, n_out: int, table=None) -> bytes:
        out = C.create_string_buffer(max(n_out, 1))
        t = (C.c_uint16 * 8)(*table) if table is not None else None
        rc = self.lib.w3o_decompress_block_ex(stream, len(stream), t, out, n_out)
        assert rc == 0, rc
        return out.raw[:n_out]

    def compress_block_geo(self, data: bytes, ctx_bits: int, align_bits: int, max_bits: int, table=None) -> bytes:
        n = len(data)
        cap = 17 * n + 64
        out = C.create_string_buffer(cap)
        ol = C.c_size_t()
        t = (C.c_uint16 * 8)(*table) if table is not None else None
        rc = self.lib.w3o_compress_block_geo(data, n, t, ctx_bits, align_bits, max_bits, out, cap, C.byref(ol))
        assert rc == 0, rc
        return out.raw[: ol.value]

    def decompress_block_geo(self, stream: bytes, n_out: int, ctx_bits: int, align_bits: int, max_bits: int, table=None) -> bytes:
        out = C.create_string_buffer(max(n_out, 1))
        t = (C.c_uint16 * 8)(*table) if table is not None else None
        rc = self.lib.w3o_decompress_block_geo(stream, len(stream), t, ctx_bits, align_bits, max_bits, out, n_out)
        assert rc == 0, rc
        return out.raw[:n_out]

    def sweep_exec(self, data: bytes, ctx_bits: int, align_bits: int, table=None) -> int:
        """bits ACStats counts for one point of the entropy-hashing-ac sweep (csize = // 8)"""
        t = (C.c_uint16 * 8)(*table) if table is not None else None
        bc = C.c_uint64()
        rc = self.lib.w3o_sweep_exec(data, len(data), t, ctx_bits, align_bits, C.byref(bc))
        assert rc == 0, rc
        return bc.value

    def cm_compress_block(self, data: bytes, log_cells: int = 16) -> bytes:
        n = len(data)
        cap = 2 * n + 64
        out = C.create_string_buffer(cap)
        ol = C.c_size_t()
        rc = self.lib.w3o_cm_compress_block(data, n, log_cells, out, cap, C.byref(ol))
        assert rc == 0, rc
        return out.raw[: ol.value]

    def cm_decompress_block(self, stream: bytes, n_out: int, log_cells: int = 16) -> bytes:
        out = C.create_string_buffer(max(n_out, 1))
        rc = self.lib.w3o_cm_decompress_block(stream, len(stream), log_cells, out, n_out)
        assert rc == 0, rc
        return out.raw[:n_out]

    def sweep_exec_ordern(self, data: bytes, ctx_bits: int, align_bits: int) -> int:
        """bits ACStats counts for one point of the ordern sweep (raw-history OrderN)"""
        bc = C.c_uint64()
        rc = self.lib.w3o_sweep_exec_ordern(data, len(data), ctx_bits, align_bits, C.byref(bc))
        assert rc == 0, rc
        return bc.value

    def stationary_new(self, data: bytes):
        t = (C.c_uint16 * 8)()
        self.lib.w3o_stationary_new(data, len(data), t)
        return tuple(t)

    def compress_file(self, data: bytes) -> bytes:
        n = len(data)
        cap = 17 * n + 64
        out = C.create_string_buffer(cap)
        ol = C.c_size_t()
        rc = self.lib.w3o_compress_file(data, n, out, cap, C.byref(ol))
        assert rc == 0, rc
        return out.raw[: ol.value]

    def decompress_file(self, blob: bytes):
        cap = max(int.from_bytes(blob[4:12], "big"), 1) if len(blob) >= 12 else 1
        out = C.create_string_buffer(cap)
        ol = C.c_size_t()
        rc = self.lib.w3o_decompress_file(blob, len(blob), out, cap, C.byref(ol))
        return rc, out.raw[: ol.value] if rc == 0 else b""

    def achash(self, bits: int, pos: int, table=None, max_bits=8):
        steps = C.c_int()
        if table is None:
            h = self.lib.w3o_achash(bits, pos, C.byref(steps))
        else:
            t = (C.c_uint16 * 8)(*table)
            h = self.lib.w3o_achash_ex(bits, pos, t, max_bits, C.byref(steps))
        return h, steps.value

    def ctx_trace(self, data: bytes):
        out = (C.c_uint16 * max(8 * len(data), 1))()
        self.lib.w3o_ctx_trace(data, len(data), out)
        return list(out[: 8 * len(data)])

    def ac_encode_probs(self, data: bytes, probs) -> bytes:
        n = len(probs)
        arr = (C.c_uint16 * max(n, 1))(*probs)
        cap = 4 * n + 16
        out = C.create_string_buffer(cap)
        ol = C.c_size_t()
        rc = self.lib.w3o_ac_encode_probs(data, len(data), arr, n, out, cap, C.byref(ol))
        assert rc == 0, rc
        return out.raw[: ol.value]

    def ac_decode_probs(self, stream: bytes, probs) -> bytes:
        n = len(probs)
        arr = (C.c_uint16 * max(n, 1))(*probs)
        cap = (n + 7) // 8
        out = C.create_string_buffer(max(cap, 1))
        rc = self.lib.w3o_ac_decode_probs(stream, len(stream), arr, n, out, cap)
        assert rc == 0, rc
        return out.raw[:cap]


class HostSim:
    """w3_core.cuh compiled for the CPU (tests/hostsim)."""

    def __init__(self, table=BOOK1, max_bits=8):
        h = C.CDLL(build_hostsim())
        sz = C.c_size_t
        h.hs_hash.restype = C.c_uint32
        h.hs_hash.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_int)]
        h.hs_hash_exact.restype = C.c_uint32
        h.hs_hash_exact.argtypes = [C.c_uint64, C.c_uint32, C.POINTER(C.c_uint16), C.c_int, C.POINTER(C.c_int)]
        h.hs_cell_count.restype = C.c_uint32
        h.hs_cell_count.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        h.hs_counter_p.restype = C.c_uint32
        h.hs_counter_p.argtypes = [C.c_uint32, C.c_uint32]
        h.hs_counter_update.restype = C.c_uint32
        h.hs_counter_update.argtypes = [C.c_uint32, C.c_uint32]
        h.hs_counter_p_sweep.restype = C.c_uint64
        h.hs_counter_p_sweep.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        h.hs_compress_block.argtypes = [C.c_void_p, C.c_char_p, sz, C.c_void_p, sz, C.POINTER(sz)]
        h.hs_decompress_block.argtypes = [C.c_void_p, C.c_char_p, sz, C.c_void_p, sz]
        h.hs_state_entry.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint16)]
        h.hs_slot_get_state.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        h.hs_slot_set_state.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        h.hs_cell_get_slot.argtypes = [C.c_void_p, C.c_uint64]
        h.hs_slot_set_nib.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint16)]
        h.hs_slot_get_nib.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint16)]
        h.hs_cell_index.restype = C.c_uint64
        h.hs_cell_index.argtypes = [C.c_uint64, C.c_uint32]
        h.hs_cm_hash.restype = C.c_uint64
        h.hs_cm_hash.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        h.hs_cm_find_slot.argtypes = [C.c_void_p, C.c_uint64]
        h.hs_cm_squash.argtypes = h.hs_cm_stretch.argtypes = [C.c_int]
        h.hs_cm_train.argtypes = [C.c_int32, C.c_int, C.c_int]
        h.hs_lane_decode_block.argtypes = [C.c_void_p, C.c_char_p, sz, C.c_uint, C.c_uint8, C.c_void_p, sz, C.c_int]
        h.hs_pair_decode_block.argtypes = [C.c_void_p, C.c_char_p, sz, C.c_uint, C.c_uint8, C.c_void_p, sz, C.c_int]
        h.hs_split_decode_block.argtypes = [C.c_void_p, C.c_char_p, sz, C.c_uint, C.c_uint8, C.c_void_p, sz, C.c_int]
        self.lib = h
        self.table = (C.c_uint16 * 8)(*table)
        self.max_bits = max_bits
        self.tab = C.create_string_buffer(h.hs_tab_size())
        self.bad = h.hs_build_table(self.table, max_bits, self.tab)

    def hash(self, bits, pos):
        slow = C.c_int()
        return self.lib.hs_hash(self.tab, bits, pos, C.byref(slow)), bool(slow.value)

    def hash_exact(self, bits, pos):
        idx = C.c_int()
        return self.lib.hs_hash_exact(bits, pos, self.table, self.max_bits, C.byref(idx)), idx.value

    def compress_block(self, data: bytes) -> bytes:
        n = len(data)
        cap = 17 * n + 64
        out = C.create_string_buffer(cap)
        ol = C.c_size_t()
        rc = self.lib.hs_compress_block(self.tab, data, n, out, cap, C.byref(ol))
        assert rc == 0
        return out.raw[: ol.value]

    def decompress_block(self, stream: bytes, n_out: int) -> bytes:
        out = C.create_string_buffer(max(n_out, 1))
        self.lib.hs_decompress_block(self.tab, stream, len(stream), out, n_out)
        return out.raw[:n_out]

    def lane_decode_block(self, stream: bytes, n_out: int, misalign: int = 0, junk: int = 0xA5, pair=False) -> bytes:
        """w3_lane.cuh's per-lane decoder (what w3_decode_blocks_lane runs in every lane), or its
        pair form (w3_decode_blocks_pair: two lanes per block in lockstep; pair="cache" = with the
        probability table), or the split form (pair="split": w3_decode_blocks_split's coder lane and
        the model lanes of a block as actors in a random order; the schedule's seed varies with the call)."""
        out = C.create_string_buffer(max(n_out, 1))
        fn = self.lib.hs_pair_decode_block if pair else self.lib.hs_lane_decode_block
        flag = 1 if pair == "cache" else 0
        if pair == "split":
            self._split_seed = getattr(self, "_split_seed", 0) + 1
            fn, flag = self.lib.hs_split_decode_block, self._split_seed
        rc = fn(self.tab, stream, len(stream), misalign, junk, out, n_out, flag)
        assert rc == 0, rc
        return out.raw[:n_out]

    def thresholds(self):
        """(thr[8][256], sing[8][256]) as python lists."""
        raw = self.tab.raw
        thr = [[int.from_bytes(raw[(a * 256 + j) * 8 : (a * 256 + j) * 8 + 8], "little") for j in range(256)] for a in range(8)]
        off = 8 * 256 * 8 + 8 * 2048 * 2 + 8 * 256
        sing = [[raw[off + a * 256 + j] for j in range(256)] for a in range(8)]
        return thr, sing


def load_all():
    return types.SimpleNamespace(oracle=Oracle(), hostsim=HostSim())
