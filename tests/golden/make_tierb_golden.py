"""Builds tests/golden/tier_b_golden.json from the reference's own dumps (run in the build
container, where /root/reference exists; the GPU box only has the committed JSON):

  docs/state_table/state_table.csv       3963 rows  state,tr0,tr1,prob   (src/state_table/naive.rs)
  docs/state_table/aux_state_table.csv    990 rows  the auxiliary table  (naive.rs:50-75)
  docs/hashslots.md:69-133                 60 rows  slot.id, bit_id, nib_ctx, rel_idx, abs_idx, parity

These are reference-issued data (the survey verified them against the source formulas), so the
Tier-B rows a17/a18 are pinned by the reference itself, unlike the Tier-A model path.
"""
import csv
import hashlib
import json
import os
import re

REF = "/root/reference/docs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tier_b_golden.json")


def table(path):
    with open(path) as f:
        rows = list(csv.DictReader(f))
    assert [int(r["state"]) for r in rows] == list(range(len(rows)))
    return {"tr0": [int(r["tr0"]) for r in rows], "tr1": [int(r["tr1"]) for r in rows],
            "prob": [int(r["prob"]) for r in rows]}


def slots(path):
    out = []
    for line in open(path).read().splitlines()[68:133]:
        cells = [c.strip() for c in line.strip().strip("|").split("|")]
        if len(cells) != 6 or not cells[0].isdigit():
            continue
        sid, bit_id, ctx, rel, ab, par = cells
        bit_id = int(re.sub(r"\D", "", bit_id))
        nib_ctx = int(re.sub(r"[^01]", "", ctx), 2)
        out.append([int(sid), bit_id, nib_ctx, int(rel), int(ab), par])
    assert len(out) == 60, len(out)
    return out


def main():
    g = {
        "source": "mitiko/weath3rb0i docs/state_table/*.csv and docs/hashslots.md:69-133",
        "sha256": {f: hashlib.sha256(open(os.path.join(REF, f), "rb").read()).hexdigest()
                   for f in ("state_table/state_table.csv", "state_table/aux_state_table.csv", "hashslots.md")},
        "state_table": table(os.path.join(REF, "state_table/state_table.csv")),
        "aux_state_table": table(os.path.join(REF, "state_table/aux_state_table.csv")),
        "hashslots": slots(os.path.join(REF, "hashslots.md")),
    }
    with open(OUT, "w") as f:
        json.dump(g, f, separators=(",", ":"))
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
