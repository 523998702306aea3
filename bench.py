#!/usr/bin/env python
"""bench.py — compress/decompress throughput of the block-parallel codec.

One step = one pass of the hot path over one batch: encode every 1 MiB block
of the batch, pack the streams, decode them again (the reference's `t` action,
src/main.rs:80-83).  Workload at N GPUs: BASELINE.json configs[1] per GPU —
64 MiB of the fixed-seed synthetic text in 1 MiB blocks — rank r coding bytes
[r*64 MiB, (r+1)*64 MiB) of the stream (weak scaling, no collective on the
data path; torch.distributed is used only for the barrier and the max).

  value  round-trip MB/s (1e6 uncompressed bytes / (encode+pack+decode time)),
         inputs resident in HBM, CUDA events on the launching stream
  e2e    the same through w3_compress / w3_decompress with pinned HOST buffers
         (H2D, kernels, D2H all inside the timed region)
  --impl reference   the CPU oracle (the reference cannot be compiled here: no
         Rust toolchain) on all host cores, one block per core at a time
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MIB = 1 << 20
METRIC = "round-trip (compress+decompress) MB/s, 64 MiB synthetic text per GPU, 1 MiB blocks"


def ncu_traffic(kernel: str, n: int):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full`
    capture (profiles/r01_ncu_metrics.json, taken on the 64 MiB bench workload);
    None for any other workload size."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_metrics.json")
    if n != 64 * MIB or not os.path.exists(p):
        return None
    with open(p) as f:
        m = json.load(f).get("decode" if "decode" in kernel else "encode", {})
    try:
        return int((float(m["dram__bytes_read.sum"]) + float(m["dram__bytes_write.sum"])) * 1e6)
    except (KeyError, ValueError):
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_round_trip(n_blocks: int, block: int, threads: int, data_ptr: int):
    """Times the CPU oracle (oracle/libw3oracle.so) on n_blocks blocks held at
    data_ptr: returns (compress_s, decompress_s, compressed_bytes)."""
    from tests import _libs

    o = _libs.load_oracle()
    n = n_blocks * block
    stride = block + block // 8 + 64
    comp = C.create_string_buffer(stride * n_blocks)
    lens = (C.c_uint32 * n_blocks)()
    back = C.create_string_buffer(n)
    t0 = time.perf_counter()
    rc = o.w3o_compress_blocks_mt(data_ptr, n, block, C.addressof(comp), stride, C.addressof(lens), threads)
    t1 = time.perf_counter()
    assert rc == 0, rc
    rc = o.w3o_decompress_blocks_mt(C.addressof(comp), stride, C.addressof(lens), C.addressof(back), n, block, threads)
    t2 = time.perf_counter()
    assert rc == 0, rc
    assert C.string_at(data_ptr, n) == back.raw[:n], "CPU oracle round trip mismatch"
    return t1 - t0, t2 - t1, sum(lens)


def workload_config(args) -> dict:
    """The `config` object, identical for both arms."""
    block = args.block_size
    n_blocks = (args.mib * MIB + block - 1) // block
    return {
        "workload": f"{args.mib} MiB synthetic text per GPU (seed 0x77333069), {block >> 10} KiB blocks, "
                    f"{n_blocks} blocks per GPU; step = compress + decompress of every block",
        "block_size": block, "blocks_per_gpu": n_blocks,
    }


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from corpus import w3_corpus
    from tests import _libs

    _libs.build_oracle()
    threads = host_threads()
    block = args.block_size
    n_blocks = max(1, min(args.mib * MIB // block, threads))  # one block per core, one round
    buf = C.create_string_buffer(n_blocks * block)
    w3_corpus.generate_into(C.addressof(buf), n_blocks * block)
    for _ in range(args.warmup):
        cpu_oracle_round_trip(min(n_blocks, threads), block, threads, C.addressof(buf))
    tc = td = 0.0
    for _ in range(args.steps):
        c, d, clen = cpu_oracle_round_trip(n_blocks, block, threads, C.addressof(buf))
        tc += c
        td += d
    nbytes = n_blocks * block * args.steps
    value = nbytes / 1e6 / (tc + td)
    sample = f"{n_blocks} x {block >> 10} KiB blocks of the same text per step, one block per core"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "MB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round((tc + td) / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": dict(workload_config(args), sample=sample),
        "compress_mbs": round(nbytes / 1e6 / tc, 3), "decompress_mbs": round(nbytes / 1e6 / td, 3),
        "cpu_baseline": {"value": round(value, 3), "unit": "MB/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist

    from corpus import w3_corpus
    from weath3rb0i_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this codec has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    block = args.block_size
    n = args.mib * MIB
    n_blocks = (n + block - 1) // block
    ctx = capi.Context([local])

    # this rank's shard of the stream, in pinned host memory
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    w3_corpus.generate_into(h_in.data_ptr(), n, offset=rank * n)
    d_in = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    d_in[:n].copy_(h_in)
    stride = (block + block // 8 + 64 + 15) // 16 * 16
    d_scratch = torch.empty(stride * n_blocks, dtype=torch.uint8, device="cuda")
    d_lens = torch.zeros(n_blocks, dtype=torch.int32, device="cuda")
    d_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_offs = torch.zeros(n_blocks + 1, dtype=torch.int64, device="cuda")
    d_packed = torch.zeros(n + n // 8 + 64, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    t_enc = t_pack = t_dec = 0.0

    def step(timed: bool):
        nonlocal t_enc, t_pack, t_dec
        e = [ev() for _ in range(4)] if timed else None
        if timed:
            e[0].record()
        ctx.encode_device(d_in.data_ptr(), n, block, d_scratch.data_ptr(), stride, d_lens.data_ptr(),
                          d_flag.data_ptr(), stream)
        if timed:
            e[1].record()
        ctx.pack_device(d_scratch.data_ptr(), stride, d_lens.data_ptr(), n_blocks, d_offs.data_ptr(),
                        d_packed.data_ptr(), stream)
        if timed:
            e[2].record()
        ctx.decode_device(d_packed.data_ptr(), d_offs.data_ptr(), n_blocks, block, n, d_out.data_ptr(), d_flag.data_ptr(), stream)
        if timed:
            e[3].record()
        return e

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    assert int(d_flag.item()) == 0, "scratch overflow or kernel abort"
    assert torch.equal(d_out[:n], d_in[:n]), "device round trip mismatch"
    comp_bytes = int(d_offs[n_blocks].item())

    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    start, end = ev(), ev()
    start.record()
    evs = [step(True) for _ in range(args.steps)]
    end.record()
    barrier()
    total_ms = start.elapsed_time(end)
    for e in evs:
        t_enc += e[0].elapsed_time(e[1])
        t_pack += e[1].elapsed_time(e[2])
        t_dec += e[2].elapsed_time(e[3])
    assert torch.equal(d_out[:n], d_in[:n]), "device round trip mismatch after timing"

    # end to end through the host-buffer C ABI, pinned host memory both sides
    cap = ctx.compress_bound(n, block)
    cap = min(cap, n + n // 4 + 4096)
    h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
    clen = 0
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        clen = ctx.compress_into(h_in.data_ptr(), n, h_comp.data_ptr(), cap, block)
        ctx.decompress_into(h_comp.data_ptr(), clen, h_back.data_ptr(), n)
    assert torch.equal(h_back, h_in), "host round trip mismatch"
    barrier()
    t0 = time.perf_counter()
    tc_e2e = 0.0
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        clen = ctx.compress_into(h_in.data_ptr(), n, h_comp.data_ptr(), cap, block)
        tc_e2e += time.perf_counter() - ta
        ctx.decompress_into(h_comp.data_ptr(), clen, h_back.data_ptr(), n)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None

    t = torch.tensor([total_ms, e2e_s * 1e3, t_enc, t_pack, t_dec, tc_e2e * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, t_enc, t_pack, t_dec, tc_e2e_ms = [float(v) for v in t.tolist()]

    if rank == 0:
        K = args.steps
        job_bytes = n * world
        value = job_bytes * K / 1e6 / (total_ms / 1e3)
        e2e_value = job_bytes * e2e_steps / 1e6 / (e2e_ms / 1e3)
        peak, peak_src = peaks()
        enc_ms, dec_ms = t_enc / K, t_dec / K
        dom = "w3_decode_blocks" if dec_ms >= enc_ms else "w3_encode_blocks"
        dom_ms = max(enc_ms, dec_ms)
        # one context lookup (one 32-byte sector) per coded bit
        alg_bytes = n * 8 * 32
        achieved = alg_bytes / 1e9 / (dom_ms / 1e3)
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965
        warps = n_blocks
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "MB/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": round(total_ms / K, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": dict(workload_config(args),
                           l2="working set per step (in + scratch + out > 190 MiB) exceeds the 126 MB L2; no explicit flush"),
            "compress_mbs": round(job_bytes / 1e6 / ((t_enc + t_pack) / K / 1e3), 3),
            "decompress_mbs": round(job_bytes / 1e6 / (dec_ms / 1e3), 3),
            "bits_per_byte": round(8.0 * comp_bytes / n, 4),
            "kernel_ms": {"w3_encode_blocks": round(enc_ms, 3), "scan+pack": round(t_pack / K, 3),
                          "w3_decode_blocks": round(dec_ms, 3)},
            "cycles_per_bit_per_warp": {"encode": round(enc_ms * 1e-3 * sm_mhz * 1e6 / (block * 8), 1),
                                        "decode": round(dec_ms * 1e-3 * sm_mhz * 1e6 / (block * 8), 1),
                                        "warps": warps},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 3), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 6),
                         "traffic": ncu_traffic(dom, n) if block == MIB else None,
                         "traffic_source": "profiles/r01_ncu_metrics.json (ncu --set full, same workload), bytes per launch",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "compulsory_hbm_bytes_per_launch": n + comp_bytes,
                         "peak_source": peak_src,
                         "algorithmic": "8 context lookups (one 32-B sector each) per input byte, SURVEY.md §8(d); "
                                        "the counters live in shared memory, so this path is latency-bound, not HBM-bound"},
            "e2e": {"value": round(e2e_value, 3), "unit": "MB/s", "h2d_bytes_per_step": n + clen,
                    "d2h_bytes_per_step": clen + 4 * n_blocks + n, "steps": e2e_steps,
                    "compress_mbs": round(job_bytes * e2e_steps / 1e6 / (tc_e2e_ms / 1e3), 3),
                    "decompress_mbs": round(job_bytes * e2e_steps / 1e6 / ((e2e_ms - tc_e2e_ms) / 1e3), 3)},
            "gpu_launches": 4 * K,
            "clocks": clocks,
        }
        # SURVEY.md §8(d), figure 2: the same 8 sectors per input byte against the box's MEASURED
        # random 32-byte-sector gather rate (w3_probe_sector_gather, outside the timed region)
        try:
            rate = ctx.sector_gather_rate(4 << 30)
            sectors = n * 8 / (dom_ms / 1e3)
            line["gather_roofline"] = {"kernel": dom, "achieved": round(sectors / 1e9, 3), "peak": round(rate / 1e9, 3),
                                       "unit": "G sectors/s", "frac": round(sectors / rate, 6),
                                       "peak_source": "measured here: one independent random 32-B sector per load over a 4 GiB buffer"}
        except Exception as e:  # a measurement aid: never fails the bench
            line["gather_roofline"] = {"error": str(e)[:120]}
        if world == 1 and not args.no_cpu:
            threads = host_threads()
            nb_cpu = max(1, min(n_blocks, threads))
            c, d, _ = cpu_oracle_round_trip(nb_cpu, block, threads, h_in.data_ptr())
            line["cpu_baseline"] = {
                "value": round(nb_cpu * block / 1e6 / (c + d), 3), "unit": "MB/s", "cores": threads, "kind": "port",
                "compress_mbs": round(nb_cpu * block / 1e6 / c, 3), "decompress_mbs": round(nb_cpu * block / 1e6 / d, 3),
                "sample": f"first {nb_cpu} x {block >> 10} KiB blocks of the same text, one block per core, round trip once"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """Everything but the final JSON line goes to stderr: libraries (NCCL's
    version banner, for one) write to fd 1 and the driver parses stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mib", type=int, default=64, help="MiB of text per GPU")
    ap.add_argument("--block-size", type=int, default=MIB)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
