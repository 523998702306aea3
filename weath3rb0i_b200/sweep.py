"""`python -m weath3rb0i_b200.sweep FILE` — the parameter sweep of the reference's research binary
src/bin/entropy-hashing-ac/main.rs on the GPU: ctx_bits 8..30 x alignment_bits 0..4, every point an
encode of the whole file into the bit counter ACStats (src/helpers.rs:60-90), all points at once
(w3_sweep, include/w3_codec.h).  Prints the lines that binary prints (main.rs:64-72, 31-36, 38-41)."""
from __future__ import annotations

import sys
import time

from . import capi


def run(path: str, ctx_range=range(8, 31), align_range=range(0, 5), out=sys.stdout, history="ac") -> dict:
    data = open(path, "rb").read()
    ctx = capi.Context([0])
    points = [(c, a) for c in ctx_range for a in align_range if a <= c]
    t0 = time.perf_counter()
    bits, table = ctx.sweep(data, points, history=history)
    tag = "eh-ac" if history == "ac" else "ordern"
    dt = time.perf_counter() - t0
    ctx.close()
    res = {p: b // 8 for p, b in zip(points, bits)}  # ACStats::result — helpers.rs:71-73
    best_all, arg_all = len(data), (0, 0)
    for c in ctx_range:
        best, arg = len(data), (0, 0)
        for a in align_range:
            if (c, a) not in res:
                continue
            size = res[(c, a)]
            print(f"[{tag}] [ctx: {c:2}, align: {a}] csize: {size} (ratio {size / max(len(data), 1):.3f})", file=out)
            if size <= best:  # `if res > best { continue }` — main.rs:27-32
                best, arg = size, (c, a)
            if size <= best_all:
                best_all, arg_all = size, (c, a)
        print(f"-> best: {best} for [ctx: {arg[0]}, align: {arg[1]}]", file=out)
    print(f"-> gloabl best: {best_all} for [ctx: {arg_all[0]}, align: {arg_all[1]}]", file=out)
    print(f"# {len(points)} points, {len(data)} bytes, {dt:.3f} s in all on the GPU; StationaryModel::new = {table}", file=out)
    return res


if __name__ == "__main__":
    if len(sys.argv) not in (2, 3) or (len(sys.argv) == 3 and sys.argv[1] not in ("eh-ac", "ordern")):
        raise SystemExit("usage: python -m weath3rb0i_b200.sweep [eh-ac|ordern] FILE   (the reference's src/bin/<name>)")
    run(sys.argv[-1], history="raw" if sys.argv[1] == "ordern" else "ac")
