"""C-ABI checks that need no GPU: the library loads, exports every symbol of
include/w3_codec.h, validates arguments, parses headers, and FAILS LOUDLY
(W3_E_CUDA) instead of falling back when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from weath3rb0i_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_lib()
    return capi.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "w3_codec.h")).read()
    declared = set(re.findall(r"\b(w3_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS)
    for s in declared:
        assert getattr(lib, s) is not None


def test_strerror_and_version(lib):
    assert lib.w3_strerror(0) == b"ok"
    assert b"magic" in lib.w3_strerror(capi.W3_E_MAGIC)
    assert b"sm_100a" in lib.w3_version()


def test_decode_launch_policy(lib):
    """The launcher's choice of decode kernel by block count (w3_decode_family): teams up to two waves of one
    CTA per SM, the split form while an SM's blocks fit one CTA of it (8), the pair form beyond; blocks too
    large for the lane decoders' 32-bit bit counters always go to the teams; options pin a family."""
    name = lambda nb, bs=1 << 20, k=0, form=0: lib.w3_decode_kernel_name(nb, bs, 148, k, form).decode()
    assert [name(nb) for nb in (1, 64, 148, 296)] == ["w3_decode_blocks_team"] * 4
    assert [name(nb) for nb in (297, 512, 1024, 1184)] == ["w3_decode_blocks_split"] * 4
    assert [name(nb) for nb in (1185, 2048, 4096)] == ["w3_decode_blocks_pair"] * 3
    assert name(4096, 32 << 20) == "w3_decode_blocks_team"  # auto and forced "many" alike
    assert name(4096, 32 << 20, 3) == "w3_decode_blocks_team"
    assert name(1024, k=1) == "w3_decode_blocks_team" and name(1024, k=2) == "w3_decode_blocks_warp"
    assert name(5, k=3) == "w3_decode_blocks_split" and name(5, k=3, form=2) == "w3_decode_blocks_pair"
    assert name(4096, k=3, form=4) == "w3_decode_blocks_split" and name(5, k=3, form=1) == "w3_decode_blocks_lane"
    assert name(5, k=3, form=3) == "w3_decode_blocks_pair"


def test_header_parsing(lib):
    ol = C.c_size_t()
    legacy = b"w30i" + (1234).to_bytes(8, "big") + b"\xff"
    assert lib.w3_decompressed_size(legacy, len(legacy), C.byref(ol)) == 0 and ol.value == 1234
    blocks = b"w3Bk\x01\0\0\0" + (5 << 20).to_bytes(8, "big") + (1 << 20).to_bytes(4, "big") + (5).to_bytes(4, "big")
    assert lib.w3_decompressed_size(blocks, len(blocks), C.byref(ol)) == 0 and ol.value == 5 << 20
    # the reference asserts on a foreign magic (src/main.rs:123-124)
    assert lib.w3_decompressed_size(b"w31i" + bytes(20), 24, C.byref(ol)) == capi.W3_E_MAGIC
    assert lib.w3_decompressed_size(b"w30i", 4, C.byref(ol)) == capi.W3_E_FORMAT
    assert lib.w3_decompressed_size(b"w3Bk\x02" + bytes(19), 24, C.byref(ol)) == capi.W3_E_FORMAT


def test_bound_and_args(lib):
    assert lib.w3_compress_bound(0, 1 << 20) >= 24
    assert lib.w3_compress_bound(1 << 20, 1 << 20) >= 17 << 20
    h = C.c_void_p()
    assert lib.w3_ctx_create(None, 0, C.byref(h)) == capi.W3_E_ARG
    assert lib.w3_ctx_create(None, 9, C.byref(h)) == capi.W3_E_ARG
    assert lib.w3_compress(None, None, 0, 0, 0, None, 0, None) == capi.W3_E_ARG


def test_no_gpu_means_error_not_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert lib.w3_ctx_create(None, 1, C.byref(h)) == capi.W3_E_CUDA
    assert not h.value
    with pytest.raises(capi.W3Error):
        capi.Context([0])
