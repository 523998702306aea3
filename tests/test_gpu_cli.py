"""The `w3` command-line driver (csrc/w3_cli.cpp) over libw3codec.so: same verbs, output naming
and messages as the reference binary (src/main.rs:24-87), plus the block container, the model
switch and windowed I/O (-w) for inputs that do not fit in host memory."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from corpus import w3_corpus  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def w3():
    from weath3rb0i_b200 import build
    return build.build_cli()


def run(w3, cwd, *args, ok=True):
    p = subprocess.run([w3, *args], cwd=cwd, capture_output=True, text=True, timeout=600)
    if ok:
        assert p.returncode == 0, p.stdout + p.stderr
    return p


def test_verbs_and_output_names(w3, tmp_path):
    """c -> <stem>.bin, d -> <stem>.orig, both in the current directory (main.rs:58-68); the
    one-block file is the reference's own format (magic w30i, main.rs:14-15)."""
    data = w3_corpus.generate(300000, offset=11)
    (tmp_path / "in").mkdir()
    (tmp_path / "in" / "book.txt").write_bytes(data)
    p = run(w3, tmp_path, "c", "in/book.txt")
    assert "Compression took:" in p.stdout  # main.rs:74
    blob = (tmp_path / "book.bin").read_bytes()
    assert blob[:4] == b"w30i" and int.from_bytes(blob[4:12], "big") == len(data)
    p = run(w3, tmp_path, "d", "book.bin")
    assert "Decompression took:" in p.stdout  # main.rs:78
    assert (tmp_path / "book.orig").read_bytes() == data
    p = run(w3, tmp_path, "t", "in/book.txt")
    assert "identical" in p.stdout


def test_directory_is_walked_one_level(w3, tmp_path):
    """main.rs:41-47"""
    d = tmp_path / "dir"
    d.mkdir()
    (d / "sub").mkdir()
    (d / "a.txt").write_bytes(w3_corpus.generate(5000, offset=1))
    (d / "b.txt").write_bytes(w3_corpus.generate(7000, offset=2))
    (d / "sub" / "c.txt").write_bytes(b"not reached")
    run(w3, tmp_path, "t", "dir")
    assert (tmp_path / "a.bin").exists() and (tmp_path / "b.bin").exists() and not (tmp_path / "c.bin").exists()


def test_usage_and_foreign_file(w3, tmp_path):
    assert run(w3, tmp_path, ok=False).returncode == 1  # main.rs:154-161
    (tmp_path / "junk.bin").write_bytes(b"this is not a w30i file at all")
    p = run(w3, tmp_path, "d", "junk.bin", ok=False)
    assert p.returncode != 0 and "magic" in (p.stdout + p.stderr).lower()


@pytest.mark.parametrize("model", ["book1", "enwik7"])
def test_windowed_io_writes_the_same_container(w3, tmp_path, model):
    """-w: a window of whole blocks at a time.  Same bytes on disk as the in-memory path, and each
    reads the other's output; the last block is ragged and the last window short."""
    data = w3_corpus.generate(5 * 65536 + 12345, offset=77)
    (tmp_path / "big.dat").write_bytes(data)
    run(w3, tmp_path, "-b", "65536", "-m", model, "c", "big.dat")
    whole = (tmp_path / "big.bin").read_bytes()
    assert whole[:4] == b"w3Bk" and whole[4] == (1 if model == "book1" else 2)
    run(w3, tmp_path, "-b", "65536", "-m", model, "-w", "131072", "c", "big.dat")
    assert (tmp_path / "big.bin").read_bytes() == whole
    run(w3, tmp_path, "-w", "131072", "d", "big.bin")
    assert (tmp_path / "big.orig").read_bytes() == data
    os.remove(tmp_path / "big.orig")
    run(w3, tmp_path, "d", "big.bin")
    assert (tmp_path / "big.orig").read_bytes() == data
    # a window smaller than a block still moves one block at a time
    run(w3, tmp_path, "-b", "65536", "-m", model, "-w", "1000", "t", "big.dat")
    assert (tmp_path / "big.bin").read_bytes() == whole


def test_windowed_training_is_refused(w3, tmp_path):
    (tmp_path / "x.dat").write_bytes(w3_corpus.generate(200000, offset=5))
    p = run(w3, tmp_path, "-b", "65536", "-m", "train", "-w", "65536", "c", "x.dat", ok=False)
    assert p.returncode != 0 and "whole input" in p.stderr


def test_reference_format_forced_above_one_mib(w3, tmp_path, libs):
    """-r 1: the reference's own one-stream file whatever the size (one block here), readable by the
    reference; without it a file above 1 MiB goes into the w3Bk container."""
    data = w3_corpus.generate((2 << 20) + 12345, offset=77)
    (tmp_path / "big.txt").write_bytes(data)
    run(w3, tmp_path, "-r", "1", "c", "big.txt")
    blob = (tmp_path / "big.bin").read_bytes()
    assert blob[:4] == b"w30i" and blob == libs.oracle.compress_file(data)
    run(w3, tmp_path, "d", "big.bin")
    assert (tmp_path / "big.orig").read_bytes() == data
    run(w3, tmp_path, "c", "big.txt")
    assert (tmp_path / "big.bin").read_bytes()[:4] == b"w3Bk"
