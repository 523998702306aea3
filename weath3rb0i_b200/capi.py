"""ctypes binding of libw3codec.so (include/w3_codec.h).

There is deliberately no fallback: if the library is missing or no CUDA device
is usable, importing succeeds but every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("W3_LIB") or os.path.join(HERE, "libw3codec.so")  # W3_LIB: another build of the same library

W3_OK, W3_E_ARG, W3_E_CAP, W3_E_MAGIC, W3_E_CUDA, W3_E_FORMAT, W3_E_TABLE, W3_E_TOOBIG = 0, -1, -2, -3, -4, -5, -6, -7
W3_OPT_KERNELS, W3_OPT_ENC_TEAMS, W3_OPT_LANE_WARPS, W3_OPT_LANE_FORM, W3_OPT_ENC_CODER, W3_OPT_SCRATCH_PERMILLE = 1, 2, 3, 4, 5, 6
KERNELS = {"auto": 0, "team": 1, "warp": 2, "many": 3}
W3_F_LEGACY = 1
DEFAULT_BLOCK = 1 << 20

# every symbol include/w3_codec.h declares
SYMBOLS = (
    "w3_ctx_create", "w3_ctx_destroy", "w3_ctx_device_count", "w3_ctx_set_option", "w3_compress_bound",
    "w3_ctx_set_model", "w3_ctx_get_model", "w3_train_model", "w3_ctx_set_geometry", "w3_ctx_get_geometry", "w3_sweep",
    "w3_compress", "w3_decompress", "w3_decompress_blocks", "w3_decompressed_size", "w3_cm_compress", "w3_cm_decompress",
    "w3_encode_device", "w3_pack_device", "w3_decode_device",
    "w3_selftest_achash", "w3_selftest_counter_p", "w3_probe_sector_gather", "w3_decode_kernel_name", "w3_selftest_ac_encode", "w3_selftest_ac_decode",
    "w3_selftest_state_table", "w3_selftest_hashslots",
    "w3_strerror", "w3_last_cuda_error", "w3_version",
)


class W3Error(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"w3 error {code}: {detail}")


class MagicMismatch(W3Error):
    """The reference asserts on a bad magic number (src/main.rs:123-124)."""


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is not built (python -m weath3rb0i_b200.build); "
            "this codec has no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    sz, vp, u8p = C.c_size_t, C.c_void_p, C.c_void_p
    lib.w3_ctx_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.w3_ctx_destroy.argtypes = [vp]
    lib.w3_ctx_destroy.restype = None
    lib.w3_ctx_device_count.argtypes = [vp]
    lib.w3_ctx_set_option.argtypes = [vp, C.c_int, C.c_int]
    lib.w3_ctx_set_model.argtypes = [vp, vp]
    lib.w3_ctx_get_model.argtypes = [vp, vp]
    lib.w3_train_model.argtypes = [vp, u8p, sz, vp]
    lib.w3_ctx_set_geometry.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.w3_ctx_get_geometry.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.w3_sweep.argtypes = [vp, C.c_int, u8p, sz, vp, vp, sz, vp, vp]
    lib.w3_compress_bound.argtypes = [sz, sz]
    lib.w3_compress_bound.restype = sz
    lib.w3_compress.argtypes = [vp, u8p, sz, sz, C.c_uint32, u8p, sz, C.POINTER(sz)]
    lib.w3_decompress.argtypes = [vp, u8p, sz, u8p, sz, C.POINTER(sz)]
    lib.w3_decompress_blocks.argtypes = [vp, u8p, sz, sz, sz, u8p, sz, C.POINTER(sz)]
    lib.w3_decompressed_size.argtypes = [u8p, sz, C.POINTER(sz)]
    lib.w3_cm_compress.argtypes = [vp, u8p, sz, sz, C.c_int, u8p, sz, C.POINTER(sz)]
    lib.w3_cm_decompress.argtypes = [vp, u8p, sz, u8p, sz, C.POINTER(sz)]
    lib.w3_encode_device.argtypes = [vp, C.c_int, vp, sz, sz, vp, sz, vp, vp, vp]
    lib.w3_pack_device.argtypes = [vp, C.c_int, vp, sz, vp, sz, vp, vp, vp]
    lib.w3_decode_device.argtypes = [vp, C.c_int, vp, vp, sz, sz, sz, vp, vp, vp]
    lib.w3_selftest_achash.argtypes = [vp, vp, vp, sz, vp]
    lib.w3_selftest_counter_p.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.w3_probe_sector_gather.argtypes = [vp, sz, C.POINTER(C.c_double)]
    lib.w3_selftest_ac_encode.argtypes = [vp, u8p, sz, vp, sz, u8p, sz, C.POINTER(sz)]
    lib.w3_selftest_ac_decode.argtypes = [vp, u8p, sz, vp, sz, u8p, sz]
    lib.w3_selftest_state_table.argtypes = [vp, C.c_int, vp, vp, vp, sz]
    lib.w3_selftest_hashslots.argtypes = [vp, vp, sz, vp, sz, vp]
    lib.w3_decode_kernel_name.argtypes = [C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int]
    for f in ("w3_strerror", "w3_last_cuda_error", "w3_version", "w3_decode_kernel_name"):
        getattr(lib, f).restype = C.c_char_p
    lib.w3_strerror.argtypes = [C.c_int]
    _lib = lib
    return lib


def check(rc: int):
    if rc == W3_OK:
        return
    lib = load()
    detail = lib.w3_strerror(rc).decode()
    if rc == W3_E_CUDA:
        detail += " — " + lib.w3_last_cuda_error().decode()
    raise (MagicMismatch if rc == W3_E_MAGIC else W3Error)(rc, detail)


def _addr(buf) -> int:
    """address of a bytes-like / ctypes / numpy / torch host buffer"""
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "data_ptr"):
        return buf.data_ptr()
    if hasattr(buf, "ctypes"):
        return buf.ctypes.data
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p).value or 0
    return C.addressof(C.c_char.from_buffer(buf)) if len(buf) else 0


class Context:
    """w3_ctx: one codec context over `devices` (ordinals) of this process."""

    def __init__(self, devices=(0,)):
        self.lib = load()
        devices = list(devices)
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        check(self.lib.w3_ctx_create(arr, len(devices), C.byref(h)))
        self.handle = h
        self.devices = devices

    def close(self):
        if getattr(self, "handle", None):
            self.lib.w3_ctx_destroy(self.handle)
            self.handle = None

    __del__ = close

    def set_option(self, option: int, value: int) -> None:
        check(self.lib.w3_ctx_set_option(self.handle, option, value))

    def set_kernels(self, name: str) -> None:
        """Pin the kernel family: auto | team | warp | many (include/w3_codec.h, W3_OPT_KERNELS)."""
        self.set_option(W3_OPT_KERNELS, KERNELS[name])

    # ---- the StationaryModel (stationary.rs:14-46) -------------------------
    def set_model(self, table) -> None:
        arr = (C.c_uint16 * 8)(*table)
        check(self.lib.w3_ctx_set_model(self.handle, C.addressof(arr)))

    def get_model(self):
        arr = (C.c_uint16 * 8)()
        check(self.lib.w3_ctx_get_model(self.handle, C.addressof(arr)))
        return tuple(arr)

    def train_model(self, data: bytes):
        """StationaryModel::new(buf) on the GPU."""
        arr = (C.c_uint16 * 8)()
        check(self.lib.w3_train_model(self.handle, _addr(data), len(data), C.addressof(arr)))
        return tuple(arr)

    # ---- the geometry of the model stack (ordern_entropy.rs:15-24) and the sweep -----------
    def set_geometry(self, bits_in_context: int, alignment_bits: int, max_bits: int) -> None:
        check(self.lib.w3_ctx_set_geometry(self.handle, bits_in_context, alignment_bits, max_bits))

    def get_geometry(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.w3_ctx_get_geometry(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def sweep(self, data, points, table=None, history="ac"):
        """exec() of src/bin/entropy-hashing-ac/main.rs:48-75 (history="ac") or compress() of
        src/bin/ordern/main.rs:66-81 (history="raw") for every (ctx_bits, alignment_bits) of
        `points`, all at once on the GPU: -> (bit counts, the StationaryModel table used)."""
        n = len(points)
        pts = (C.c_uint8 * (2 * n))(*[v for p in points for v in p])
        t = (C.c_uint16 * 8)(*table) if table is not None else None
        used = (C.c_uint16 * 8)()
        out = (C.c_uint64 * n)()
        check(self.lib.w3_sweep(self.handle, {"ac": 0, "raw": 1}[history], _addr(data), len(data), C.addressof(t) if t is not None else None,
                                C.addressof(pts), n, C.addressof(out), C.addressof(used)))
        return list(out), tuple(used)

    # ---- host-buffer API -------------------------------------------------
    def compress_bound(self, n: int, block_size: int = DEFAULT_BLOCK) -> int:
        return self.lib.w3_compress_bound(n, block_size)

    def compress_into(self, src, n: int, dst, cap: int, block_size: int = DEFAULT_BLOCK, legacy=False) -> int:
        ol = C.c_size_t()
        check(self.lib.w3_compress(self.handle, _addr(src), n, block_size,
                                   W3_F_LEGACY if legacy else 0, _addr(dst), cap, C.byref(ol)))
        return ol.value

    def compress(self, data: bytes, block_size: int = DEFAULT_BLOCK, legacy=False) -> bytes:
        n = len(data)
        cap = max(n + n // 4 + 4096, 4096)
        while True:
            out = C.create_string_buffer(cap)
            ol = C.c_size_t()
            rc = self.lib.w3_compress(self.handle, _addr(data), n, block_size,
                                      W3_F_LEGACY if legacy else 0, C.addressof(out), cap, C.byref(ol))
            if rc == W3_E_CAP and ol.value > cap:
                cap = ol.value
                continue
            check(rc)
            return out.raw[: ol.value]

    def decompressed_size(self, blob: bytes) -> int:
        ol = C.c_size_t()
        check(self.lib.w3_decompressed_size(_addr(blob), len(blob), C.byref(ol)))
        return ol.value

    def decompress_into(self, src, n_in: int, dst, cap: int) -> int:
        ol = C.c_size_t()
        check(self.lib.w3_decompress(self.handle, _addr(src), n_in, _addr(dst), cap, C.byref(ol)))
        return ol.value

    def decompress_blocks_into(self, src, n_in: int, first_block: int, n_blocks: int, dst, cap: int) -> int:
        ol = C.c_size_t()
        check(self.lib.w3_decompress_blocks(self.handle, _addr(src), n_in, first_block, n_blocks, _addr(dst), cap, C.byref(ol)))
        return ol.value

    def decompress_blocks(self, blob: bytes, first_block: int, n_blocks: int, block_size: int) -> bytes:
        cap = max(n_blocks * block_size, 1)
        out = C.create_string_buffer(cap)
        n = self.decompress_blocks_into(blob, len(blob), first_block, n_blocks, C.addressof(out), cap)
        return out.raw[:n]

    def decompress(self, blob: bytes) -> bytes:
        n = self.decompressed_size(blob)
        out = C.create_string_buffer(max(n, 1))
        ol = C.c_size_t()
        check(self.lib.w3_decompress(self.handle, _addr(blob), len(blob), C.addressof(out), n, C.byref(ol)))
        return out.raw[: ol.value]

    # ---- the hashed context-mixing stack (include/w3_codec.h, w3_cm_*) ---------------------
    def cm_compress(self, data: bytes, block_size: int = DEFAULT_BLOCK, log_cells: int = 16) -> bytes:
        n = len(data)
        cap = n + n // 4 + 4096 + 4 * (n // block_size + 1)
        out = C.create_string_buffer(cap)
        ol = C.c_size_t()
        check(self.lib.w3_cm_compress(self.handle, _addr(data), n, block_size, log_cells, C.addressof(out), cap, C.byref(ol)))
        return out.raw[: ol.value]

    def cm_decompress(self, blob: bytes) -> bytes:
        n = int.from_bytes(blob[8:16], "big") if len(blob) >= 16 else 0
        out = C.create_string_buffer(max(n, 1))
        ol = C.c_size_t()
        check(self.lib.w3_cm_decompress(self.handle, _addr(blob), len(blob), C.addressof(out), n, C.byref(ol)))
        return out.raw[: ol.value]

    # ---- device-resident API (pointers are ints; stream is a cudaStream_t int)
    def encode_device(self, d_in, n, block_size, d_scratch, stride, d_lens, d_overflow, stream=0, dev=0):
        check(self.lib.w3_encode_device(self.handle, dev, d_in, n, block_size, d_scratch, stride,
                                        d_lens, d_overflow, stream))

    def pack_device(self, d_scratch, stride, d_lens, n_blocks, d_offsets, d_packed, stream=0, dev=0):
        check(self.lib.w3_pack_device(self.handle, dev, d_scratch, stride, d_lens, n_blocks,
                                      d_offsets, d_packed, stream))

    def decode_device(self, d_comp, d_offsets, n_blocks, block_size, n, d_out, d_err=None, stream=0, dev=0):
        check(self.lib.w3_decode_device(self.handle, dev, d_comp, d_offsets, n_blocks, block_size,
                                        n, d_out, d_err, stream))

    # ---- stage self-tests ------------------------------------------------
    def achash(self, bits, pos):
        n = len(bits)
        b = (C.c_uint64 * max(n, 1))(*bits)
        p = (C.c_uint32 * max(n, 1))(*pos)
        o = (C.c_uint32 * max(n, 1))()
        check(self.lib.w3_selftest_achash(self.handle, C.addressof(b), C.addressof(p), n, C.addressof(o)))
        return list(o[:n])

    def counter_p_mismatches(self) -> int:
        m = C.c_uint64(1)
        check(self.lib.w3_selftest_counter_p(self.handle, C.byref(m)))
        return m.value

    def sector_gather_rate(self, buffer_bytes: int = 4 << 30) -> float:
        """Random 32-byte-sector loads per second of the first device (gather roofline denominator)."""
        r = C.c_double(0.0)
        check(self.lib.w3_probe_sector_gather(self.handle, buffer_bytes, C.byref(r)))
        return r.value

    def state_table(self, which: int = 0):
        """(next0, next1, prob) lists of the device-built NaiveStateTable (0) / auxiliary table (1)."""
        n = 3963 if which == 0 else 990
        a, b, p = ((C.c_uint16 * n)() for _ in range(3))
        check(self.lib.w3_selftest_state_table(self.handle, which, C.addressof(a), C.addressof(b), C.addressof(p), n))
        return list(a), list(b), list(p)

    def hashslot_ops(self, ops, cells: bytearray):
        """Replays cell operations [(code, cell, a, b, c), ...] on the device; returns the result
        pairs and updates `cells` (n x 96 bytes) in place."""
        n = len(ops)
        flat = (C.c_uint32 * (4 * max(n, 1)))()
        for i, (code, cell, a, b, c) in enumerate(ops):
            flat[4 * i:4 * i + 4] = [code | (cell << 8), a, b, c]
        buf = (C.c_uint8 * len(cells)).from_buffer(cells)
        res = (C.c_uint32 * (2 * max(n, 1)))()
        check(self.lib.w3_selftest_hashslots(self.handle, C.addressof(flat), n, C.addressof(buf), len(cells) // 96, C.addressof(res)))
        return [(res[2 * i], res[2 * i + 1]) for i in range(n)]

    def ac_encode_probs(self, data: bytes, probs) -> bytes:
        n = len(probs)
        arr = (C.c_uint16 * max(n, 1))(*probs)
        cap = 4 * n + 16
        out = C.create_string_buffer(cap)
        ol = C.c_size_t()
        check(self.lib.w3_selftest_ac_encode(self.handle, _addr(data), len(data), C.addressof(arr), n,
                                             C.addressof(out), cap, C.byref(ol)))
        return out.raw[: ol.value]

    def ac_decode_probs(self, stream: bytes, probs) -> bytes:
        n = len(probs)
        arr = (C.c_uint16 * max(n, 1))(*probs)
        cap = (n + 7) // 8
        out = C.create_string_buffer(max(cap, 1))
        check(self.lib.w3_selftest_ac_decode(self.handle, _addr(stream), len(stream), C.addressof(arr), n,
                                             C.addressof(out), cap))
        return out.raw[:cap]
