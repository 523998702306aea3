"""Extract the judged metrics from the two `ncu --set full` captures of scripts/gpu_profiles.sh
(gpurun_out/prof_encode.ncu-rep, prof_decode.ncu-rep) into profiles/r01_ncu_metrics.json, the
launch list into profiles/r01_launches_bench.csv, and print the tables of profiles/README.md.
Run here (ncu reads reports without a GPU):  python scripts/extract_ncu.py"""
import collections, csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["short_scoreboard", "wait", "branch_resolving", "no_instruction", "long_scoreboard", "barrier",
          "math_pipe_throttle", "mio_throttle", "dispatch_stall", "lg_throttle", "not_selected", "selected",
          "sleeping", "membar"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    res = {}
    for name in ("encode", "decode"):
        rep = os.path.join(OUT, f"prof_{name}.ncu-rep")
        if not os.path.exists(rep):
            print("missing", rep); continue
        m = raw(rep)
        d = {}
        for k in KEEP + [STALL % s for s in STALLS]:
            if k in m:
                v, u = m[k]
                v = v.replace(",", "")
                if k.startswith("dram__bytes") and u.lower().startswith("g"):  # keep MB
                    v = str(float(v) * 1e3)
                if k.startswith("dram__bytes") and u.lower().startswith("k"):
                    v = str(float(v) / 1e3)
                d[k] = v
                d.setdefault("_units", {})[k] = u
        res[name] = d
    with open(os.path.join(ROOT, "profiles", "r01_ncu_metrics.json"), "w") as f:
        json.dump(res, f, indent=1)
    lc = os.path.join(OUT, "launches.csv")
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(ROOT, "profiles", "r01_launches_bench.csv"))
        agg = collections.OrderedDict()
        with open(lc) as f:
            lines = [l for l in f if l.startswith('"')]
        for r in csv.DictReader(io.StringIO("".join(lines))):
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            u = r["Metric Unit"]
            ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
            k = r["Kernel Name"][:48]
            a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
        tot = sum(a[1] for a in agg.values())
        print("| kernel | launches | avg ms | share |\n|---|---|---|---|")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| `{k}` | {c} | {t / c:.3f} | {100 * t / tot:.2f} % |")
    print()
    if len(res) == 2:
        e, d = res["encode"], res["decode"]
        print("| metric | `w3_encode_blocks_team` | `w3_decode_blocks_team` |\n|---|---|---|")
        for k in KEEP + [STALL % s for s in STALLS]:
            if k in e or k in d:
                print(f"| `{k}` ({e.get('_units', {}).get(k, '')}) | {e.get(k, '')} | {d.get(k, '')} |")


if __name__ == "__main__":
    main()
