"""N > 1 host path on CPU: world_size-2 gloo ranks each code a contiguous block
range and rank 0 joins the containers; the result must be independent of the
number of ranks.  The per-block coder is the oracle here (no GPU in this
test): what is under test is the sharding arithmetic, the container assembly
and the torch.distributed plumbing bench.py relies on."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from corpus import w3_corpus
from weath3rb0i_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, n, block, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests._libs import Oracle

    oracle = Oracle()
    lo, hi = sharding.byte_range(n, block, world, rank)
    data = w3_corpus.generate(hi - lo, offset=lo)
    streams = [oracle.compress_block(data[i : i + block]) for i in range(0, len(data), block)]
    part = sharding.build_container(len(data), block, streams)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(part, gathered, dst=0)
    dist.barrier()
    if rank == 0:
        with open(out_path, "wb") as f:
            f.write(sharding.join_containers(gathered))
    dist.destroy_process_group()


def _rank_in_place(rank, world, port, n, block, out_path):
    """bench.py's e2e join: exchange only the stream lengths, every rank writes its streams at the
    offsets the index implies into ONE shared file, rank 0 adds header and index."""
    import numpy as np
    import torch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests._libs import Oracle

    oracle = Oracle()
    nb = sharding.n_blocks_of(n, block)
    lo, hi = sharding.byte_range(n, block, world, rank)
    data = w3_corpus.generate(hi - lo, offset=lo)
    streams = [oracle.compress_block(data[i : i + block]) for i in range(0, len(data), block)]
    per = sharding.lens_per_rank(nb, world)
    mine = torch.zeros(per, dtype=torch.int32)
    mine[: len(streams)] = torch.tensor([len(s) for s in streams], dtype=torch.int32)
    gathered = [torch.zeros(per, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, mine)
    all_lens = sharding.flatten_gathered_lens([g.tolist() for g in gathered], nb)
    start, end, total = sharding.placement(n, block, world, rank, all_lens)
    if rank == 0:
        with open(out_path, "wb") as f:
            f.truncate(total)
    dist.barrier()
    shm = np.memmap(out_path, dtype=np.uint8, mode="r+", shape=(total,))
    body = b"".join(streams)
    assert end - start == len(body)
    shm[start:end] = np.frombuffer(body, dtype=np.uint8)
    if rank == 0:
        head = sharding.header_bytes(n, block, nb) + sharding.index_bytes(all_lens)
        shm[: len(head)] = np.frombuffer(head, dtype=np.uint8)
    shm.flush()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,block", [(65536 * 5 + 777, 65536), (4096 * 7 + 1, 4096)])
def test_two_ranks_join_in_place(tmp_path, libs, n, block):
    out = str(tmp_path / "shared.w3bk")
    mp.spawn(_rank_in_place, args=(2, _free_port(), n, block, out), nprocs=2, join=True)
    data = w3_corpus.generate(n)
    single = sharding.build_container(n, block, [libs.oracle.compress_block(data[i : i + block]) for i in range(0, n, block)])
    assert open(out, "rb").read() == single


@pytest.mark.parametrize("n,block", [(65536 * 5 + 777, 65536), (4096 * 3, 4096), (1000, 4096)])
def test_two_ranks_equal_one(tmp_path, libs, n, block):
    out = str(tmp_path / "joined.bin")
    mp.spawn(_rank_main, args=(2, _free_port(), n, block, out), nprocs=2, join=True)
    joined = open(out, "rb").read()
    data = w3_corpus.generate(n)
    single = sharding.build_container(n, block, [libs.oracle.compress_block(data[i : i + block]) for i in range(0, n, block)])
    assert joined == single
    nn, bs, streams = sharding.split_container(joined)
    assert (nn, bs) == (n, block)
    back = b"".join(libs.oracle.decompress_block(s, min(block, n - i * block)) for i, s in enumerate(streams))
    assert back == data


def test_block_ranges_cover_everything():
    for nb in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                b0, b1 = sharding.block_range(nb, world, r)
                assert 0 <= b0 <= b1 <= nb
                got += list(range(b0, b1))
            assert got == list(range(nb))


def test_container_round_trip_and_errors():
    streams = [b"abc", b"", b"\xff" * 10]
    blob = sharding.build_container(2 * 8 + 3, 8, streams)
    assert sharding.split_container(blob) == (19, 8, streams)
    assert sharding.offsets_of(streams) == [0, 3, 3, 13]
    with pytest.raises(ValueError):
        sharding.split_container(blob[:-1])
    with pytest.raises(ValueError):
        sharding.split_container(b"w30i" + blob[4:])
    with pytest.raises(ValueError):
        sharding.join_containers([sharding.build_container(5, 8, [b"x"]), sharding.build_container(8, 8, [b"y"])])
