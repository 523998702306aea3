"""Extract the judged metrics from the round-2 `ncu --set full` captures (scripts/gpu_r2_profiles.sh:
gpurun_out/r02_prof_*.ncu-rep) into profiles/r02_ncu_metrics.json (what bench.py reads for `traffic`
and `issue_roofline`), the launch list into profiles/r02_launches_bench.csv with a per-kernel summary,
and print the tables of profiles/README.md.  Run here (ncu reads reports without a GPU)."""
import collections, csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum", "dram__sectors_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum",
]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["short_scoreboard", "wait", "branch_resolving", "no_instruction", "long_scoreboard", "barrier",
          "math_pipe_throttle", "mio_throttle", "dispatch_stall", "lg_throttle", "not_selected", "selected", "sleeping", "membar"]
UNIT = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    return [{h: (v, u) for h, u, v in zip(hdr, units, r)} for r in rows[2:]]


def pick(m):
    d = {"kernel": m.get("Kernel Name", ("", ""))[0]}
    for k in KEEP + [STALL % s for s in STALLS]:
        if k in m:
            v, u = m[k]
            v = v.replace(",", "")
            try:
                f = float(v)
            except ValueError:
                continue
            if k.startswith("dram__bytes"):
                f *= UNIT.get(u.lower(), 1)
            if k == "gpu__time_duration.sum":  # -> milliseconds
                f *= {"second": 1e3, "s": 1e3, "msecond": 1.0, "ms": 1.0, "usecond": 1e-3, "us": 1e-3, "nsecond": 1e-6, "ns": 1e-6}.get(u.lower(), 1.0)
            d[k] = f
    return d


def main():
    res, bench = {}, {}
    reps = {"decode_split": "r02_prof_decode_split", "decode_pair": "r02_prof_decode_pair", "encode_team": "r02_prof_encode_team",
            "gather_probe": "r02_prof_gather_probe", "hbm_paths": "r02_prof_hbm_paths"}
    prev = {}
    try:  # captures made earlier in the round stay as they are when their reports are no longer here
        prev = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")))
    except (OSError, ValueError):
        pass
    for name, f in reps.items():
        rep = os.path.join(OUT, f + ".ncu-rep")
        if not os.path.exists(rep):
            print("missing", rep)
            if name in prev.get("details", {}):
                res[name] = prev["details"][name]
            continue
        rows = [pick(m) for m in raw_rows(rep)]
        res[name] = rows if len(rows) > 1 else (rows[0] if rows else {})
    # what bench.py wants: per-kernel DRAM bytes and warp instructions per launch on the 1 GiB workload
    n = 1 << 30
    for key, kn in (("decode_split", "w3_decode_blocks_split"), ("decode_pair", "w3_decode_blocks_pair"), ("encode_team", "w3_encode_blocks_team")):
        m = res.get(key)
        if isinstance(m, dict) and m:
            bench[kn] = {"input_bytes": n,
                         "dram_bytes": m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0),
                         "warp_instructions": m.get("smsp__inst_executed.sum", 0),
                         "duration_ms_under_ncu": m.get("gpu__time_duration.sum", 0)}
    bench["details"] = res
    with open(os.path.join(ROOT, "profiles", "r02_ncu_metrics.json"), "w") as f:
        json.dump(bench, f, indent=1)
    lc = os.path.join(OUT, "r02_launches.csv")
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(ROOT, "profiles", "r02_launches_bench.csv"))
        agg = collections.OrderedDict()
        with open(lc) as f:
            lines = [l for l in f if l.startswith('"')]
        for r in csv.DictReader(io.StringIO("".join(lines))):
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            k = r["Kernel Name"][:60]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)
        tot = sum(v[1] for v in agg.values())
        print("| kernel | launches | avg ms | share |\n|---|---|---|---|")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| `{k}` | {c} | {ms / c:.3f} | {100 * ms / tot:.2f} % |")
    for name, m in res.items():
        for row in (m if isinstance(m, list) else [m]):
            print("\n##", name, row.get("kernel", "")[:70])
            for k, v in row.items():
                if k != "kernel":
                    print(f"  {k}: {v:.6g}")


if __name__ == "__main__":
    main()
