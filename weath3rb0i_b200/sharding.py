"""Host-side block sharding and container assembly (no GPU needed).

Blocks are independent, so N GPUs (or N ranks) each take a contiguous range of
blocks and the host concatenates the per-block streams and rebuilds the offset
index — no collective on the data path.  The same arithmetic
lives in w3_capi.cu (block_range) for the in-process multi-device path; this
module is what bench.py and multi-process callers use.

Container (include/w3_codec.h):
  "w3Bk" | u8 version | u8[3] 0 | u64 BE n | u32 BE block_size | u32 BE n_blocks
  | version 2 only: 8 x u16 BE StationaryModel probabilities
  | n_blocks x u32 BE stream length | streams back to back
(this module builds and joins version-1 containers, i.e. the for_book1 model)
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

MAGIC_BLOCKS = b"w3Bk"
MAGIC_LEGACY = b"w30i"  # src/main.rs:14
HDR_BLOCKS = 24


def n_blocks_of(n: int, block_size: int) -> int:
    return (n + block_size - 1) // block_size


def block_range(n_blocks: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block range [b0, b1) of `rank` out of `world` (ceil split)."""
    per = (n_blocks + world - 1) // world
    b0 = min(n_blocks, per * rank)
    return b0, min(n_blocks, b0 + per)


def byte_range(n: int, block_size: int, world: int, rank: int) -> Tuple[int, int]:
    b0, b1 = block_range(n_blocks_of(n, block_size), world, rank)
    return min(n, b0 * block_size), min(n, b1 * block_size)


def split_container(blob: bytes):
    """-> (n, block_size, [stream bytes per block])."""
    if blob[:4] != MAGIC_BLOCKS or blob[4] != 1:
        raise ValueError("not a w3Bk v1 container")
    n = int.from_bytes(blob[8:16], "big")
    bs = int.from_bytes(blob[16:20], "big")
    nb = int.from_bytes(blob[20:24], "big")
    if nb != n_blocks_of(n, bs) or len(blob) < HDR_BLOCKS + 4 * nb:
        raise ValueError("inconsistent container header")
    lens = [int.from_bytes(blob[HDR_BLOCKS + 4 * i : HDR_BLOCKS + 4 * i + 4], "big") for i in range(nb)]
    at, streams = HDR_BLOCKS + 4 * nb, []
    for ln in lens:
        streams.append(blob[at : at + ln])
        at += ln
    if at != len(blob):
        raise ValueError("container length does not match its index")
    return n, bs, streams


def build_container(n: int, block_size: int, streams: Sequence[bytes]) -> bytes:
    if len(streams) != n_blocks_of(n, block_size):
        raise ValueError("stream count does not match n / block_size")
    head = MAGIC_BLOCKS + bytes([1, 0, 0, 0]) + n.to_bytes(8, "big") + block_size.to_bytes(4, "big") + len(streams).to_bytes(4, "big")
    index = b"".join(len(s).to_bytes(4, "big") for s in streams)
    return head + index + b"".join(streams)


def join_containers(parts: Sequence[bytes]) -> bytes:
    """Concatenate per-rank containers (rank order) into one.  Every part but
    the last must hold whole blocks of the same block size."""
    n_total, bs, streams = 0, None, []  # type: int, int | None, List[bytes]
    for i, p in enumerate(parts):
        n, b, s = split_container(p)
        if not s:
            continue
        if bs is None:
            bs = b
        if b != bs:
            raise ValueError("parts use different block sizes")
        if n_total % bs:
            raise ValueError(f"part {i} follows a ragged block")
        n_total += n
        streams += s
    if bs is None:
        _, bs, _ = split_container(parts[0])
    return build_container(n_total, bs, streams)


def offsets_of(streams: Sequence[bytes]) -> List[int]:
    """Exclusive prefix sum of stream lengths: the offset index a decoder uses."""
    out, run = [0], 0
    for s in streams:
        run += len(s)
        out.append(run)
    return out


# ---- joining the ranks' streams in place (bench.py's e2e leg; tests/test_sharding_gloo.py) --------
# The only thing the ranks of a multi-process job exchange is the per-block stream length (the
# offset index, 4 bytes per block): from it every rank knows where its streams go in the ONE
# container, and writes them there itself (a shared mapping, or pwrite into a shared file).


def lens_per_rank(n_blocks: int, world: int) -> int:
    """Slots per rank of the padded length array that an all_gather exchanges."""
    return (n_blocks + world - 1) // world


def flatten_gathered_lens(gathered: Sequence[Sequence[int]], n_blocks: int) -> List[int]:
    """gathered[r][:] = rank r's padded lengths -> the index of the whole container."""
    world = len(gathered)
    out: List[int] = []
    for r in range(world):
        b0, b1 = block_range(n_blocks, world, r)
        out += [int(v) for v in gathered[r][: b1 - b0]]
    return out


def header_bytes(n: int, block_size: int, n_blocks: int) -> bytes:
    return MAGIC_BLOCKS + bytes([1, 0, 0, 0]) + n.to_bytes(8, "big") + block_size.to_bytes(4, "big") + n_blocks.to_bytes(4, "big")


def index_bytes(all_lens: Sequence[int]) -> bytes:
    return b"".join(int(v).to_bytes(4, "big") for v in all_lens)


def placement(n: int, block_size: int, world: int, rank: int, all_lens: Sequence[int]) -> Tuple[int, int, int]:
    """-> (byte offset of this rank's first stream in the container, one past its last, container length)."""
    nb = n_blocks_of(n, block_size)
    if len(all_lens) != nb:
        raise ValueError("index length does not match n / block_size")
    offs = [0]
    for v in all_lens:
        offs.append(offs[-1] + int(v))
    hdr = HDR_BLOCKS + 4 * nb
    b0, b1 = block_range(nb, world, rank)
    return hdr + offs[b0], hdr + offs[b1], hdr + offs[nb]
