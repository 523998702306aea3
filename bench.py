#!/usr/bin/env python
"""bench.py — compress/decompress throughput of the block-parallel codec.

Workload (BASELINE.json configs[2]): ONE input — the first
`--mib` MiB (default 1024) of the fixed-seed synthetic text — in 1 MiB blocks, sharded over the
N GPUs as contiguous block ranges (rank r: sharding.block_range(n_blocks, N, r)); strong scaling.
One step = one pass of the hot path over the whole input: encode every block, pack the streams,
decode them again (the reference's `t` action, src/main.rs:80-83).  No collective on the data
path: torch.distributed carries the barrier, the max over ranks and (e2e) the per-block stream
lengths — 4 bytes per block — from which every rank derives where its streams go.

  value  round-trip MB/s (1e6 uncompressed bytes of the WHOLE input / (encode+pack+decode time)),
         inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e    the same through the C ABI with HOST buffers: every rank w3_compress()es its shard from
         pinned host memory (H2D, kernels, D2H inside), the ranks join their streams and the offset
         index into ONE w3Bk container in shared host memory (/dev/shm; rank 0 writes header and
         index), and every rank w3_decompress_blocks()es its block range of that container back.
         The container's SHA-256 must not depend on N (asserted against a constant for the
         default workload).
  config2  (N = 1 only) BASELINE.json configs[1], 64 MiB in 64 blocks: last round's headline
  --impl reference   the CPU oracle (the reference is Rust and cannot be compiled here: no
         toolchain) on all host cores, one block per core at a time, on a bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MIB = 1 << 20
# SHA-256 of the w3Bk container of the default workload (1 GiB, 1 MiB blocks): the same at every N
CONTAINER_SHA256 = {(1024, MIB): "8491ae22b3f6d6d0d4f66b9a86f127c6f93bb96044058007389020a7fd3eff35"}  # 479 042 365 bytes


def metric_name(mib: int, block: int) -> str:
    size = f"{mib // 1024} GiB" if mib % 1024 == 0 else f"{mib} MiB"
    return f"round-trip (compress+decompress) MB/s, {size} synthetic text, {block >> 10} KiB blocks sharded over the GPUs"


def ncu_metrics():
    """Per-kernel figures from the committed `ncu --set full` captures of this workload
    (profiles/r02_ncu_metrics.json): DRAM bytes and warp instructions per launch."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        return json.load(f)


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
        "workload": f"BASELINE.json configs[2]: {args.mib} MiB synthetic text (seed 0x77333069), {block >> 10} KiB blocks, "
                    f"{n_blocks} blocks in all, contiguous block ranges over the GPUs; step = compress + decompress of every block",
        "block_size": block, "n_blocks": n_blocks, "input_bytes": args.mib * MIB,
    }


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_blocks(n_blocks: int, threads: int) -> int:
    """Blocks of the workload the CPU arm codes per step: one per host thread (a bounded sample —
    the whole 1 GiB would take the oracle more than a minute per step)."""
    return max(1, min(n_blocks, threads))


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
    n_blocks = cpu_sample_blocks(args.mib * MIB // block, threads)
    buf = C.create_string_buffer(n_blocks * block)
    w3_corpus.generate_into(C.addressof(buf), n_blocks * block)
    for _ in range(args.warmup):
        cpu_oracle_round_trip(n_blocks, block, threads, C.addressof(buf))
    tc = td = 0.0
    for _ in range(args.steps):
        c, d, clen = cpu_oracle_round_trip(n_blocks, block, threads, C.addressof(buf))
        tc += c
        td += d
    nbytes = n_blocks * block * args.steps
    value = nbytes / 1e6 / (tc + td)
    sample = f"the first {n_blocks} x {block >> 10} KiB blocks of the workload per step, one block per core, MB/s = bytes coded / time"
    line = {
        "impl": "reference", "metric": metric_name(args.mib, block), "value": round(value, 3), "unit": "MB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round((tc + td) / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args),
        "compress_mbs": round(nbytes / 1e6 / tc, 3), "decompress_mbs": round(nbytes / 1e6 / td, 3),
        "cpu_baseline": {"value": round(value, 3), "unit": "MB/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


class DeviceLeg:
    """encode + pack + decode of `n` device-resident bytes in `block`-byte blocks, timed with CUDA events."""

    def __init__(self, torch, ctx, h_in, n, block):
        self.torch, self.ctx, self.n, self.block = torch, ctx, n, block
        self.nb = (n + block - 1) // block
        self.d_in = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
        self.d_in[:n].copy_(h_in[:n])
        self.stride = (block + block // 8 + 64 + 15) // 16 * 16
        self.d_scratch = torch.empty(self.stride * max(self.nb, 1), dtype=torch.uint8, device="cuda")
        self.d_lens = torch.zeros(max(self.nb, 1), dtype=torch.int32, device="cuda")
        self.d_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.d_offs = torch.zeros(self.nb + 1, dtype=torch.int64, device="cuda")
        self.d_packed = torch.zeros(n + n // 8 + 64 * (self.nb + 1), dtype=torch.uint8, device="cuda")
        self.d_out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
        self.stream = torch.cuda.current_stream().cuda_stream

    def step(self, timed: bool):
        t, c = self.torch, self.ctx
        e = [t.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None
        if timed:
            e[0].record()
        c.encode_device(self.d_in.data_ptr(), self.n, self.block, self.d_scratch.data_ptr(), self.stride,
                        self.d_lens.data_ptr(), self.d_flag.data_ptr(), self.stream)
        if timed:
            e[1].record()
        c.pack_device(self.d_scratch.data_ptr(), self.stride, self.d_lens.data_ptr(), self.nb, self.d_offs.data_ptr(),
                      self.d_packed.data_ptr(), self.stream)
        if timed:
            e[2].record()
        c.decode_device(self.d_packed.data_ptr(), self.d_offs.data_ptr(), self.nb, self.block, self.n,
                        self.d_out.data_ptr(), self.d_flag.data_ptr(), self.stream)
        if timed:
            e[3].record()
        return e

    def check(self, what: str):
        assert int(self.d_flag.item()) == 0, f"scratch overflow or kernel abort ({what})"
        assert self.torch.equal(self.d_out[: self.n], self.d_in[: self.n]), f"device round trip mismatch ({what})"

    def comp_bytes(self) -> int:
        return int(self.d_offs[self.nb].item())


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from corpus import w3_corpus
    from weath3rb0i_b200 import capi, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this codec has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    block = args.block_size
    n_total = args.mib * MIB
    nb_total = sharding.n_blocks_of(n_total, block)
    b0, b1 = sharding.block_range(nb_total, world, rank)
    lo, hi = sharding.byte_range(n_total, block, world, rank)
    n, nb = hi - lo, b1 - b0
    ctx = capi.Context([local])

    # this rank's shard of the ONE input, in pinned host memory
    h_in = torch.empty(max(n, 1), dtype=torch.uint8).pin_memory()
    if n:
        w3_corpus.generate_into(h_in.data_ptr(), n, offset=lo)
    leg = DeviceLeg(torch, ctx, h_in, n, block)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        leg.step(False)
    barrier()
    leg.check("warm-up")
    comp_local = leg.comp_bytes()

    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    evs = [leg.step(True) for _ in range(args.steps)]
    end.record()
    barrier()
    total_ms = start.elapsed_time(end)
    t_enc = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_pack = sum(e[1].elapsed_time(e[2]) for e in evs)
    t_dec = sum(e[2].elapsed_time(e[3]) for e in evs)
    leg.check("after timing")

    # ---- end to end: host buffers through the C ABI, ONE container joined in shared host memory ----
    hdr_total = sharding.HDR_BLOCKS + 4 * nb_total
    shm_cap = hdr_total + n_total + n_total // 8 + 64 * (nb_total + 1)
    shm_dir = "/dev/shm"
    try:  # shared host memory for the joined container; a plain temporary directory if /dev/shm is too small
        import shutil
        if shutil.disk_usage(shm_dir).free < shm_cap + (64 << 20):
            raise OSError("small /dev/shm")
    except OSError:
        import tempfile
        shm_dir = tempfile.gettempdir()
    shm_path = os.path.join(shm_dir, f"w3_bench_{os.environ.get('MASTER_PORT', 'solo')}_{os.getuid()}.w3bk")
    if rank == 0:
        with open(shm_path, "wb") as f:
            f.truncate(shm_cap)
    barrier()
    shm = np.memmap(shm_path, dtype=np.uint8, mode="r+", shape=(shm_cap,))
    pinned_shm = False
    try:  # page-lock the shared mapping so that the decode leg's H2D runs at pinned speed
        pinned_shm = int(torch.cuda.cudart().cudaHostRegister(shm.ctypes.data, shm_cap, 0)) == 0
    except Exception:
        pinned_shm = False
    cap = min(ctx.compress_bound(n, block), n + n // 4 + 4096 + 8 * nb)
    h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_back = torch.empty(max(n, 1), dtype=torch.uint8).pin_memory()
    comp_np = h_comp.numpy()
    per = sharding.lens_per_rank(nb_total, world)  # padded index slice per rank
    d_mylens = torch.zeros(per, dtype=torch.int32, device="cuda")
    d_alllens = torch.zeros(per * world, dtype=torch.int32, device="cuda")
    state = {"total": 0}

    def e2e_step():
        """-> seconds spent in the compress + join half"""
        ta = time.perf_counter()
        clen = ctx.compress_into(h_in.data_ptr(), n, h_comp.data_ptr(), cap, block) if n else sharding.HDR_BLOCKS
        lens = comp_np[sharding.HDR_BLOCKS: sharding.HDR_BLOCKS + 4 * nb].view(">u4").astype(np.int64)
        if world > 1:  # the offset index is the only thing the ranks exchange: 4 bytes per block
            d_mylens.zero_()
            d_mylens[:nb].copy_(torch.from_numpy(lens.astype(np.int32)))
            dist.all_gather_into_tensor(d_alllens, d_mylens)
            all_lens = np.asarray(sharding.flatten_gathered_lens(d_alllens.cpu().numpy().reshape(world, per), nb_total), dtype=np.int64)
        else:
            all_lens = lens
        start, end, total = sharding.placement(n_total, block, world, rank, all_lens)
        shm[start:end] = comp_np[sharding.HDR_BLOCKS + 4 * nb: clen]  # this rank's streams, at their place
        if rank == 0:
            shm[: sharding.HDR_BLOCKS] = np.frombuffer(sharding.header_bytes(n_total, block, nb_total), dtype=np.uint8)
            shm[sharding.HDR_BLOCKS: hdr_total] = all_lens.astype(">u4").view(np.uint8)
        state["total"] = total
        barrier()  # the container is complete
        tb = time.perf_counter()
        if n:
            got = ctx.decompress_blocks_into(shm.ctypes.data, state["total"], b0, nb, h_back.data_ptr(), n)
            assert got == n
        return tb - ta

    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    barrier()
    assert torch.equal(h_back[:n], h_in[:n]), "host round trip mismatch"
    t0 = time.perf_counter()
    tc_e2e = 0.0
    for _ in range(e2e_steps):
        tc_e2e += e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    assert torch.equal(h_back[:n], h_in[:n]), "host round trip mismatch after timing"
    total_len = state["total"]

    red = torch.tensor([total_ms, e2e_s * 1e3, t_enc, t_pack, t_dec, tc_e2e * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([comp_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, t_enc, t_pack, t_dec, tc_e2e_ms = [float(v) for v in red.tolist()]
    comp_bytes = int(tot.item())

    if rank == 0:
        K = args.steps
        sha = hashlib.sha256(shm[:total_len]).hexdigest()
        want = CONTAINER_SHA256.get((args.mib, block))
        assert want is None or sha == want, f"container SHA-256 {sha} differs from the recorded one: the result depends on N"
        value = n_total * K / 1e6 / (total_ms / 1e3)
        e2e_value = n_total * e2e_steps / 1e6 / (e2e_ms / 1e3)
        peak, peak_src = peaks()
        enc_ms, dec_ms = t_enc / K, t_dec / K
        decode_dom = dec_ms >= enc_ms
        dom_ms = max(enc_ms, dec_ms)
        if decode_dom:  # the launcher's own choice: teams up to two waves of one CTA per SM, the split form up to 8 blocks per SM, lane pairs beyond
            try:
                sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
                dom = ctx.lib.w3_decode_kernel_name(nb, block, sms, 0, 0).decode()
            except Exception:  # naming the kernel never fails the bench: the same policy, restated
                dom = "w3_decode_blocks_team" if nb <= 2 * 148 else "w3_decode_blocks_split" if -(-nb // 148) <= 8 else "w3_decode_blocks_pair"
        else:
            dom = "w3_encode_blocks_team"
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965
        prof = ncu_metrics().get(dom, {})
        # SURVEY.md §8(d) figure 2: one context lookup = one 32-byte sector per coded bit, against the
        # device's MEASURED random-sector load rate (w3_probe_sector_gather, outside the timed region)
        sectors = n * 8  # per launch of the dominant kernel on this rank
        alg_bytes = sectors * 32
        achieved = alg_bytes / 1e9 / (dom_ms / 1e3)
        try:
            rate = ctx.sector_gather_rate(4 << 30)
            gather_peak = rate * 32 / 1e9
            gerr = None
        except Exception as e:  # a measurement aid: never fails the bench
            gather_peak, gerr = None, str(e)[:120]
        traffic = None
        if prof.get("dram_bytes") and prof.get("input_bytes"):
            traffic = int(prof["dram_bytes"] * n / prof["input_bytes"])
        line = {
            "metric": metric_name(args.mib, block), "value": round(value, 3), "unit": "MB/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": round(total_ms / K, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": dict(workload_config(args)),
            "l2": "working set per step (input + scratch + packed + output) exceeds the 126 MB L2 at every N; no explicit flush",
            "blocks_per_gpu": nb,
            "compress_mbs": round(n_total / 1e6 / ((t_enc + t_pack) / K / 1e3), 3),
            "decompress_mbs": round(n_total / 1e6 / (dec_ms / 1e3), 3),
            "bits_per_byte": round(8.0 * comp_bytes / n_total, 4),
            "container_sha256": sha, "container_sha256_checked": want is not None,
            "kernel_ms": {"w3_encode_blocks": round(enc_ms, 3), "scan+pack": round(t_pack / K, 3),
                          "w3_decode_blocks": round(dec_ms, 3)},
            "sm_cycles_per_bit": {"encode": round(enc_ms * 1e-3 * sm_mhz * 1e6 * 148 / (n * 8), 2),
                                  "decode": round(dec_ms * 1e-3 * sm_mhz * 1e6 * 148 / (n * 8), 2),
                                  "note": "SM-cycles of the whole GPU per coded bit of this rank's blocks"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 3),
                         "peak": round(gather_peak, 3) if gather_peak else None,
                         "unit": "GB/s", "frac": round(achieved / gather_peak, 6) if gather_peak else None,
                         "traffic": traffic,
                         "traffic_source": "profiles/r02_ncu_metrics.json (ncu --set full of this kernel on this workload), scaled to this rank's bytes",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "compulsory_hbm_bytes_per_launch": n + comp_local,
                         "peak_source": "measured here: random 32-B-sector loads over a 4 GiB buffer (w3_probe_sector_gather) x 32 B — "
                                        "the gather roofline of SURVEY.md §8(d).2" if gather_peak else gerr,
                         "algorithmic": "8 context lookups (one 32-B sector each) per input byte, SURVEY.md §8(d); the counters "
                                        "live in shared memory, so this path is bound by instruction issue and latency, not by HBM"},
            "hbm_roofline": {"achieved": round(achieved, 3), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 6),
                             "peak_source": peak_src},
            "e2e": {"value": round(e2e_value, 3), "unit": "MB/s", "h2d_bytes_per_step": n_total + comp_bytes,
                    "d2h_bytes_per_step": comp_bytes + 4 * nb_total + n_total, "steps": e2e_steps,
                    "compress_mbs": round(n_total * e2e_steps / 1e6 / (tc_e2e_ms / 1e3), 3),
                    "decompress_mbs": round(n_total * e2e_steps / 1e6 / max((e2e_ms - tc_e2e_ms) / 1e3, 1e-9), 3),
                    "join": "per-rank streams + offset index joined into one w3Bk container in shared host memory, inside the timed region",
                    "container_bytes": total_len, "shared_mapping_pinned": pinned_shm},
            "gpu_launches": 4 * K * world,
            "clocks": clocks,
        }
        if prof.get("warp_instructions") and prof.get("input_bytes"):
            inst = prof["warp_instructions"] * n / prof["input_bytes"]
            line["issue_roofline"] = {
                "kernel": dom, "warp_instructions_per_bit": round(prof["warp_instructions"] / (prof["input_bytes"] * 8), 2),
                "achieved": round(inst / (dom_ms / 1e3) / 1e9, 3), "peak": round(148 * 4 * sm_mhz * 1e6 / 1e9, 3),
                "unit": "G warp-instructions/s", "frac": round(inst / (dom_ms / 1e3) / (148 * 4 * sm_mhz * 1e6), 4),
                "note": "instructions per launch from the ncu capture x this rank's share, over 148 SMs x 4 schedulers x SM clock"}
        if world == 1 and not args.no_config2 and n_total > 64 * MIB and block == MIB:
            # BASELINE.json configs[1] (last round's headline workload), for continuity
            leg2 = DeviceLeg(torch, ctx, h_in, 64 * MIB, MIB)
            for _ in range(2):
                leg2.step(False)
            torch.cuda.synchronize()
            ev2 = [leg2.step(True) for _ in range(2)]
            torch.cuda.synchronize()
            leg2.check("config2")
            e2, d2 = (sum(e[0].elapsed_time(e[2]) for e in ev2) / 2, sum(e[2].elapsed_time(e[3]) for e in ev2) / 2)
            line["config2"] = {"workload": "BASELINE.json configs[1]: 64 MiB, 1 MiB blocks, 64 blocks on one GPU (one block per SM)",
                               "value": round(64 * MIB / 1e6 / ((e2 + d2) / 1e3), 3), "unit": "MB/s",
                               "compress_mbs": round(64 * MIB / 1e6 / (e2 / 1e3), 3), "decompress_mbs": round(64 * MIB / 1e6 / (d2 / 1e3), 3),
                               "cycles_per_bit_per_block": {"encode": round(e2 * 1e-3 * sm_mhz * 1e6 / (MIB * 8), 1),
                                                            "decode": round(d2 * 1e-3 * sm_mhz * 1e6 / (MIB * 8), 1)}}
            del leg2
        if world == 1 and not args.no_cpu:
            threads = host_threads()
            nb_cpu = cpu_sample_blocks(nb_total, threads)
            c, d, _ = cpu_oracle_round_trip(nb_cpu, block, threads, h_in.data_ptr())
            line["cpu_baseline"] = {
                "value": round(nb_cpu * block / 1e6 / (c + d), 3), "unit": "MB/s", "cores": threads, "kind": "port",
                "compress_mbs": round(nb_cpu * block / 1e6 / c, 3), "decompress_mbs": round(nb_cpu * block / 1e6 / d, 3),
                "sample": f"the first {nb_cpu} x {block >> 10} KiB blocks of the workload, one block per core, round trip once"}
        emit(line)
    barrier()
    if pinned_shm:
        try:
            torch.cuda.cudart().cudaHostUnregister(shm.ctypes.data)
        except Exception:
            pass
    del shm
    if rank == 0:
        try:
            os.unlink(shm_path)
        except OSError:
            pass
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
    ap.add_argument("--mib", type=int, default=1024, help="MiB of text in the whole job (sharded over the GPUs)")
    ap.add_argument("--block-size", type=int, default=MIB)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-config2", action="store_true", help="skip the 64 MiB sub-measurement")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
