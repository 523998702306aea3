"""File-level mirror of the reference driver (src/main.rs:24-144).

Same names, argument meaning and output naming as the reference:
  compress(input_file, output_file)    src/main.rs:89
  decompress(input_file, output_file)  src/main.rs:115
  run(path, action)                    src/main.rs:55   (c → <stem>.bin, d → <stem>.orig, in the CWD)
`t` additionally compares the round trip with the input (the reference does not).
"""
from __future__ import annotations

import enum
import os
import sys
import time

from . import capi

_default = None


def default_context() -> capi.Context:
    global _default
    if _default is None:
        n = int(os.environ.get("W3_DEVICES", "1"))
        _default = capi.Context(range(n))
    return _default


class Action(enum.Enum):  # src/main.rs:17-22
    Compress = "c"
    Decompress = "d"
    Test = "t"


BOOK1 = (1, 50188, 62497, 15819, 22545, 31499, 22988, 29616)  # stationary.rs:40-42
ENWIK7 = (752, 50314, 58928, 21421, 24680, 30788, 24297, 32530)  # stationary.rs:44-46


def compress(input_file, output_file, block_size: int | None = None, ctx: capi.Context | None = None,
             model="book1") -> None:
    """src/main.rs:89-113.  block_size=None writes the reference's own
    single-stream file (magic w30i) when the input fits one block, otherwise
    the w3Bk block container with 1 MiB blocks.  model: "book1" (init_model,
    src/main.rs:151), "enwik7", "train" (StationaryModel::new on the input,
    stationary.rs:14-34) or 8 probabilities; anything but book1 writes the
    container, which carries the model."""
    ctx = ctx or default_context()
    with open(input_file, "rb") as f:
        data = f.read()
    table = {"book1": BOOK1, "enwik7": ENWIK7}.get(model, model) if model != "train" else ctx.train_model(data)
    ctx.set_model(table)
    if tuple(table) != BOOK1:
        block_size = block_size or capi.DEFAULT_BLOCK
    if block_size is None:
        legacy = len(data) <= capi.DEFAULT_BLOCK
        blob = ctx.compress(data, capi.DEFAULT_BLOCK, legacy=legacy)
    else:
        blob = ctx.compress(data, block_size)
    with open(output_file, "wb") as f:
        f.write(blob)


def decompress(input_file, output_file, ctx: capi.Context | None = None) -> None:
    """src/main.rs:115-144; raises MagicMismatch where the reference asserts (:124)."""
    ctx = ctx or default_context()
    with open(input_file, "rb") as f:
        blob = f.read()
    data = ctx.decompress(blob)
    with open(output_file, "wb") as f:
        f.write(data)


def run(path, action: Action, block_size: int | None = None, out=sys.stdout) -> bool:
    """src/main.rs:55-87: outputs land in the current directory."""
    stem = os.path.splitext(os.path.basename(path))[0]

    def _c(p):
        dst = os.path.join(os.getcwd(), stem + ".bin")  # main.rs:58-62
        t = time.perf_counter()
        compress(p, dst, block_size)
        print(f"Compression took: {time.perf_counter() - t:.6f}s", file=out)  # main.rs:74
        return dst

    def _d(p):
        dst = os.path.join(os.getcwd(), stem + ".orig")  # main.rs:64-68
        t = time.perf_counter()
        decompress(p, dst)
        print(f"Decompression took: {time.perf_counter() - t:.6f}s", file=out)  # main.rs:78
        return dst

    if action is Action.Compress:
        _c(path)
    elif action is Action.Decompress:
        _d(path)
    else:  # main.rs:80-83, plus the comparison helpers::cmp would do (helpers.rs:7-28)
        orig = _d(_c(path))
        with open(path, "rb") as a, open(orig, "rb") as b:
            same = a.read() == b.read()
        print("round trip: " + ("identical" if same else "MISMATCH"), file=out)
        return same
    return True


def main(argv=None) -> int:
    """weath3rb0i <c|d|t> <path> — src/main.rs:24-53, usage at 154-161."""
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2 or argv[0] not in ("c", "d", "t"):
        print("Usage: weath3rb0i <action> <path>\n  action: c (compress) | d (decompress) | t (test)",
              file=sys.stderr)
        return 1
    action = Action(argv[0])
    path = argv[1]
    ok = True
    if os.path.isdir(path):  # main.rs:41-47 — one level, files only
        for name in sorted(os.listdir(path)):
            p = os.path.join(path, name)
            if os.path.isfile(p):
                ok &= run(p, action)
    else:
        ok = run(path, action)
    return 0 if ok else 2
