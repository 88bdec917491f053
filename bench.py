#!/usr/bin/env python
"""bench.py -- graph-build throughput (k-mer occurrences/s) of the CUDA path, with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A step = one whole graph build (parse -> k-mer extract -> hash aggregate -> Node records) of one batch of
synthetic reads. N=1 runs BASELINE.json configs[1] (E. coli-sized genome, 50x, 150 bp, 1 % error, k=31);
N>1 is weak scaling: every rank builds from its own cfg2-sized shard of reads drawn from an N-times larger
genome, k-mers are hash-partitioned to their owner GPU (NCCL all-to-all-v) before insertion.
`value`  : device-timed, reads resident in HBM -> records resident in HBM (CUDA events, max over ranks).
`e2e`    : the same build through the public host API with pinned HOST buffers, H2D of the text and D2H of
           the whole record stream inside the timed region (wall clock bracketed by synchronises).
`--impl reference`: the reference's algorithm on the host CPU cores (oracle/gx_oracle.c, a C port: the
           reference itself is Java and no JVM exists here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def kb_of(k):
    return 8 * ((k + 31) // 32)


def algorithmic_bytes(k, L, n_occ, n_distinct):
    """bytes = N*B_occ + D*B_dist (whole job) and the insert kernel's share N*B_occ + D*Kb."""
    kb, nb = kb_of(k), (k + 3) // 4
    b_occ = L / (L - k + 1) + kb + 16
    b_dist = kb + (kb + 8) + (4 + nb) + (1 + 2 * (8 + nb) + 4)
    return n_occ * b_occ + n_distinct * b_dist, n_occ * b_occ + n_distinct * kb, b_occ, b_dist


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(name: str, rank: int, world: int):
    from genomix_b200 import synth
    base = synth.CONFIGS[name]
    if world == 1:
        w = base
        cache = os.environ.get("GX_BENCH_TEXT_CACHE")
        if cache and os.path.exists(cache + f".{name}.npy"):
            return w, np.load(cache + f".{name}.npy"), base.n_reads
        text = synth.readid_text(w)
        if cache:
            np.save(cache + f".{name}.npy", text)
        return w, text, base.n_reads
    # weak scaling: genome x world, this rank's shard = base.n_reads reads with globally unique ids
    w = synth.Workload(f"{name}x{world}", base.genome_bp * world, base.read_len, base.coverage, base.error, base.k,
                       base.paired, base.seed, base.outer_mean, base.outer_std)
    return w, synth.shard_text(w, rank, base.n_reads), base.n_reads


def expected_distinct(w, n_reads):
    """capacity hint a driver can derive from its own options (genome size estimate + error rate)"""
    mates = 2 if w.paired else 1
    err_kmers = n_reads * mates * w.read_len * w.error * min(w.k, w.read_len - w.k + 1)
    return int(min(w.genome_bp + err_kmers, n_reads * mates * (w.read_len - w.k + 1)) * 1.05) + 1024


def run_reference(args, rank, world):
    """CPU arm: the reference's pipeline shape (C port) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    from genomix_b200 import synth
    from oracle import c_oracle
    c_oracle.load()
    base = synth.CONFIGS[args.workload]
    cores = os.cpu_count() or 1
    n_sample = min(base.n_reads, args.ref_sample_reads)
    text = synth.readid_text(base, n_reads=n_sample)
    occ = synth.occurrences(base, n_sample)
    for _ in range(args.warmup):
        c_oracle.build_graph_records(base.k, text, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.build_graph_records(base.k, text, cores)
    dt = time.perf_counter() - t0
    value = occ * args.steps / dt
    sample = f"first {n_sample} reads of {args.workload} ({occ} k-mer occurrences) per step"
    line = {
        "impl": "reference", "metric": "kmer_occurrences_per_sec_graph_build", "value": value, "unit": "kmers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_desc(base, 1), "k": base.k, "read_len": base.read_len, "sample": sample},
        "bases_per_sec": n_sample * base.read_len * (2 if base.paired else 1) * args.steps / dt,
        "cpu_baseline": {"value": value, "unit": "kmers/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Java (no JVM in the image): C port of its pipeline (parse, sort, group, hash shuffle, group), oracle/gx_oracle.c",
    }
    print(json.dumps(line), flush=True)


def workload_desc(w, world):
    s = (f"{w.name}: {w.genome_bp * world} bp random genome, {w.coverage:g}x, {w.read_len} bp "
         f"{'paired' if w.paired else 'single-end'} reads, {w.error * 100:g}% substitutions, k={w.k}")
    return s


def line_aligned_chunks(text_np, chunk_bytes):
    """[(offset, length)] of whole-line chunks of at most chunk_bytes (host-side; the pushes themselves are device-side)."""
    n = int(text_np.size)
    out, pos = [], 0
    while pos < n:
        end = min(n, pos + chunk_bytes)
        if end < n:
            back = np.flatnonzero(text_np[max(pos, end - (1 << 16)):end] == 10)
            if back.size:
                end = max(pos, end - (1 << 16)) + int(back[-1]) + 1
        out.append((pos, end - pos))
        pos = end
    return out


class Job:
    """One workload resident on this rank's GPU plus the GraphBuilder that builds it."""

    def __init__(self, gx, torch, dist, dev, local_rank, rank, world, w, text_np, hint, uid=None, chunk_bytes=256 << 20,
                 stream_records=False, share=None):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev
        self.w = w
        if share is not None:   # same texts as another job (no second copy)
            self.host_text, self.dev_text = share.host_text, share.dev_text
        else:
            self.host_text = torch.from_numpy(text_np).pin_memory()
            self.dev_text = self.host_text.to(dev, non_blocking=False)
        self.gb = gx.GraphBuilder(w.k, device=local_rank, rank=rank, n_ranks=world, expected_kmers=hint, chunk_bytes=chunk_bytes,
                                  stream_records=stream_records)
        self.stream = torch.cuda.current_stream(dev)
        self.gb.set_stream(self.stream.cuda_stream)
        self.chunks = line_aligned_chunks(text_np, chunk_bytes) if (world > 1 or stream_records) else [(0, int(text_np.size))]
        self.rounds = len(self.chunks)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                uid = torch.from_numpy(self.gb.mg_unique_id().copy())
            uid = uid.to(dev)
            dist.broadcast(uid, 0)
            self.gb.mg_init(uid.cpu().numpy())
            t = torch.tensor([self.rounds], device=dev, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            self.rounds = int(t.item())

    def step_device(self):
        """reads resident in HBM -> records resident in HBM; N > 1: one collective exchange round per chunk"""
        gb = self.gb
        gb.reset()
        base = self.dev_text.data_ptr()
        for i in range(self.rounds):
            if i < len(self.chunks):
                off, ln = self.chunks[i]
                gb.push_lines_device(base + off, ln)
            if self.world > 1:
                gb.mg_exchange()
        gb.finish()

    def step_host(self, out_host):
        """pinned host text -> records in pinned host memory (H2D and D2H inside)"""
        import ctypes as C
        gb = self.gb
        gb.reset()
        if self.world == 1:
            gb.push_lines(self.host_text)   # the library cuts it into chunks and double-buffers the H2D copies
        for i in range(self.rounds if self.world > 1 else 0):
            if i < len(self.chunks):
                off, ln = self.chunks[i]
                gb.push_lines(self.host_text[off:off + ln])
            gb.mg_exchange()
        gb.finish()
        n = gb.record_bytes
        cursor, used, pos = C.c_uint64(0), C.c_size_t(0), 0
        while pos < n:
            gb._check(gb._lib.gx_next_records(gb._ctx, C.byref(cursor), C.c_void_p(out_host.data_ptr() + pos),
                                               out_host.numel() - pos, C.byref(used)))
            if used.value == 0:
                break
            pos += used.value
        return pos

    def step_device_streamed(self, sink):
        """reads resident in HBM -> records serialised slice by slice and drained into a small pinned host buffer (for
        workloads whose record stream does not fit in HBM next to the table, e.g. BASELINE configs[4])"""
        import ctypes as C
        gb = self.gb
        gb.reset()
        base = self.dev_text.data_ptr()
        for i in range(self.rounds):
            if i < len(self.chunks):
                off, ln = self.chunks[i]
                gb.push_lines_device(base + off, ln)
            if self.world > 1:
                gb.mg_exchange()
        gb.finish()
        n = gb.record_bytes
        cursor, used, pos = C.c_uint64(0), C.c_size_t(0), 0
        while pos < n:
            gb._check(gb._lib.gx_next_records(gb._ctx, C.byref(cursor), C.c_void_p(sink.data_ptr()), sink.numel(), C.byref(used)))
            if used.value == 0:
                break
            pos += used.value
        return pos

    def timed_streamed(self, steps, warmup):
        torch = self.torch
        sink = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        for _ in range(warmup):
            self.step_device_streamed(sink)
        stats = self.gb.stats()
        self.barrier()
        t0 = time.perf_counter()
        phase_acc = {}
        for _ in range(steps):
            self.step_device_streamed(sink)
            for key, val in self.gb.phase_ms().items():
                phase_acc[key] = phase_acc.get(key, 0.0) + val
        self.barrier()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, {k_: v / steps for k_, v in phase_acc.items()}, 0, stats

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, steps, warmup):
        """(ms per step, max over ranks; per-phase ms per step; launches per step; stats)"""
        torch = self.torch
        for _ in range(warmup):
            self.step_device()
        stats = self.gb.stats()
        self.barrier()
        launches0 = self.gb.kernel_launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phase_acc = {}
        ev0.record(self.stream)
        for _ in range(steps):
            self.step_device()
            for key, val in self.gb.phase_ms().items():
                phase_acc[key] = phase_acc.get(key, 0.0) + val
        ev1.record(self.stream)
        self.barrier()
        ms = ev0.elapsed_time(ev1) / steps
        launches = (self.gb.kernel_launches - launches0) / steps
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, {k_: v / steps for k_, v in phase_acc.items()}, launches, stats

    def allsum(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor([float(v) for v in vals], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def close(self):
        self.gb.close()
        del self.dev_text, self.host_text


def parity_pass(gx, torch, dist, dev, local_rank, rank, world, name, factor):
    """Untimed multi-GPU parity check: a 1/factor twin of `name` is built by all ranks from line shards of the same text;
    the ranks' order-independent canonical fingerprints (oracle/gx_oracle.c, additive over disjoint record sets) are
    summed and compared with the C oracle's fingerprint of the whole text on rank 0."""
    from oracle import c_oracle as CO
    from genomix_b200 import multigpu
    w = gx.synth.scaled(gx.synth.CONFIGS[name], factor)
    text = gx.synth.readid_text(w)
    shard = multigpu.shard_lines(text, rank, world)
    job = Job(gx, torch, dist, dev, local_rank, rank, world, w, np.ascontiguousarray(shard), 0, chunk_bytes=max(1 << 20, shard.size // 3 + 1))
    job.step_device()
    stream = np.frombuffer(job.gb.records(), dtype=np.uint8)
    fp = CO.canonical_fingerprint(stream)
    st = job.gb.stats()
    job.close()
    ints = torch.tensor([fp.sum & 0xffffffff, fp.sum >> 32, fp.records, fp.heads, fp.edges, int(fp.coverage_total), st["kmer_occurrences"]],
                        device=dev, dtype=torch.int64)
    gathered = [torch.zeros_like(ints) for _ in range(world)]
    dist.all_gather(gathered, ints)
    xors = torch.tensor([fp.xor & 0x7fffffffffffffff, fp.xor >> 63], device=dev, dtype=torch.int64)
    xg = [torch.zeros_like(xors) for _ in range(world)]
    dist.all_gather(xg, xors)
    out = None
    if rank == 0:
        tot = [0] * 7
        x = 0
        for g, xx in zip(gathered, xg):
            v = [int(t) for t in g.tolist()]
            tot = [a + b for a, b in zip(tot, v)]
            lo, hi = (int(t) for t in xx.tolist())
            x ^= lo | (hi << 63)
        got_sum = ((tot[0] + (tot[1] << 32)) & 0xffffffffffffffff)
        want_stream, ost = CO.build_graph_records(w.k, text, os.cpu_count() or 1, as_numpy=True)
        wfp = CO.canonical_fingerprint(want_stream)
        ok = (got_sum == wfp.sum and x == wfp.xor and tot[2] == wfp.records and tot[3] == wfp.heads and tot[4] == wfp.edges
              and tot[5] == int(wfp.coverage_total) and tot[6] == ost["occurrences"])
        out = {"ok": bool(ok), "workload": f"{name}/{factor:g} (k={w.k}, {w.n_reads} reads, {ost['occurrences']} k-mer occurrences)",
               "records": tot[2], "oracle_records": int(wfp.records), "read_heads": tot[3], "edges": tot[4], "n_ranks": world,
               "check": "sum/xor of per-record 64-bit hashes (key, neighbour sets, read-head lists, coverage) over all ranks == C oracle"}
    flag = torch.tensor([1 if (out is None or out["ok"]) else 0], device=dev)
    dist.broadcast(flag, 0)
    return out, bool(flag.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--target-workload", default="cfg4", help="N>1: second, strong-scaled full-size workload (BASELINE configs[3]); 'none' skips it")
    ap.add_argument("--target-steps", type=int, default=2)
    ap.add_argument("--target-stream", action="store_true", help="target workload: stream the records to the host instead of keeping them in HBM")
    ap.add_argument("--ref-sample-reads", type=int, default=60000)
    ap.add_argument("--cpu-sample-reads", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-nohint", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import genomix_b200 as gx

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; genomix_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- multi-GPU parity, untimed, before anything is measured: k=55 twin of the target workload
    parity = None
    if world > 1 and not args.no_parity:
        parity, ok = parity_pass(gx, torch, dist, dev, local_rank, rank, world, "cfg4", 100)
        if not ok:
            if rank == 0:
                print(json.dumps({"metric": "kmer_occurrences_per_sec_graph_build", "value": None, "parity": parity}), flush=True)
            dist.destroy_process_group()
            raise SystemExit("bench.py: multi-GPU parity check FAILED")

    w, text_np, n_reads = make_workload(args.workload, rank, world)
    n_occ = gx.synth.occurrences(w, n_reads)
    n_bases = n_reads * w.read_len * (2 if w.paired else 1)
    hint = expected_distinct(w, n_reads * world) // world
    # N > 1: 64 MiB chunks, one exchange round each, so that a chunk's NVLink transfer runs under the next chunk's split
    job = Job(gx, torch, dist, dev, local_rank, rank, world, w, text_np, hint, chunk_bytes=(64 << 20) if world > 1 else (256 << 20))
    gb = job.gb

    sampler = ClockSampler(local_rank)
    sampler.start()  # started before the warm-up so that short timed regions still get samples (all under load)
    ms_per_step, phase, launches, stats = job.timed(args.steps, args.warmup)
    clocks = sampler.stop()
    tot_occ, tot_bases, tot_distinct, launches = job.allsum(n_occ, n_bases, stats["distinct_kmers"], launches)
    value = tot_occ / (ms_per_step * 1e-3)

    # ---- roofline, timed live with CUDA events inside the library (gx_phase_ms), per rank 0
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    kb = kb_of(w.k)
    n_up = stats["distinct_kmers"]            # keys this rank's table ends up with
    rank_occ = tot_occ / world                  # records this rank upserts (own + received; hash-uniform)
    job_bytes, build_bytes, b_occ, b_dist = algorithmic_bytes(w.k, w.read_len, n_occ, n_up)
    upsert_ms = phase.get("insert", 0.0) + phase.get("exchange_insert", 0.0)
    build_ms = upsert_ms + phase.get("split", 0.0)
    upsert_bytes = rank_occ * (kb + 16) + n_up * kb     # per occurrence: its key + the 16-byte slot update; per new key: the key
    # DRAM traffic of the kernels, from the committed ncu --set full capture of this workload (profiles/traffic.json, written
    # by scripts/ncu_to_profiles.py: one entry per launch of one step; cold-cache and serialised, so bytes, not times)
    traffic, traffic_src, kernel_traffic = None, None, {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and os.path.exists(tpath):
        try:
            tj = json.load(open(tpath)).get(f"{args.workload}_kernels")
            if tj:
                traffic_src = tj["source"]
                for kname, launches_ in tj["launches"].items():
                    kernel_traffic[kname] = {"dram_bytes_per_step": sum(x["dram_bytes"] for x in launches_), "launches": len(launches_),
                                             "ncu_ms": sum(x["ms"] for x in launches_), "files": [x["file"] for x in launches_]}
                up = [v for k_, v in kernel_traffic.items() if k_.startswith("upsert_regions_kernel")]
                traffic = sum(v["dram_bytes_per_step"] for v in up) if up else None
        except Exception:
            traffic, kernel_traffic = None, {}
    # algorithmic bytes of the other two phases (BASELINE.md §3 split by phase): split reads each base once for the placement and
    # every 16th line once more for the bucket sizes, and writes one (Kb + 2)-byte record per occurrence; finish reads every
    # slot once and writes the records
    split_bytes = n_occ * ((1 + 1 / 16) * w.read_len / (w.read_len - w.k + 1) + kb + 2)
    finish_bytes = stats["table_capacity"] * {8: 16, 16: 32, 24: 32, 32: 48}[kb] + stats["record_bytes"] if "record_bytes" in stats else None
    roof = {
        "bound": "hbm", "kernel": "upsert_regions_kernel<KW> (region-sorted k-mer records -> hash-table upserts)",
        "achieved": upsert_bytes / (upsert_ms * 1e-3) / 1e9 if upsert_ms > 0 else None,
        "peak": peak, "unit": "GB/s", "peak_source": peak_src,
        "frac": (upsert_bytes / (upsert_ms * 1e-3) / 1e9 / peak) if upsert_ms > 0 else None,
        "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_step": upsert_bytes, "kernel_ms_per_step": upsert_ms,
        "launches_per_step": max(1, job.rounds),
        "bytes_per_occurrence": kb + 16,
        "build_phase": {"kernels": "split_count + split_place + upsert_regions (what round 1's fused extract_kernel did)",
                        "algorithmic_bytes_per_step": build_bytes, "ms_per_step": build_ms, "bytes_per_occurrence": b_occ,
                        "frac": build_bytes / (build_ms * 1e-3) / 1e9 / peak if build_ms > 0 else None},
        "phases": {
            "split": {"kernels": "split_count + split_place", "algorithmic_bytes_per_step": split_bytes, "ms_per_step": phase.get("split", 0.0),
                      "frac": split_bytes / (phase["split"] * 1e-3) / 1e9 / peak if phase.get("split") else None},
            "finish": {"kernels": "heads_* + emit_scan + emit_write", "algorithmic_bytes_per_step": finish_bytes,
                       "ms_per_step": phase.get("finish", 0.0),
                       "frac": finish_bytes / (phase["finish"] * 1e-3) / 1e9 / peak if (finish_bytes and phase.get("finish")) else None},
        },
        "kernel_traffic": kernel_traffic or None,
        "job_bytes_per_step": job_bytes,
        "job_frac": job_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
    }
    nvlink = None
    if world > 1:
        comm_ms = phase.get("exchange_comm", 0.0)
        sent = stats["exchanged_records"] * (kb + 2)   # exact: records this rank pushed to other ranks in one step
        nvlink = {"bytes_sent_per_rank_per_step": sent, "exchange_comm_ms": comm_ms,
                  "achieved_GBps_per_direction": sent / (comm_ms * 1e-3) / 1e9 if comm_ms > 0 else None,
                  "peak_GBps_per_direction": 900.0, "frac": sent / (comm_ms * 1e-3) / 1e9 / 900.0 if comm_ms > 0 else None,
                  "measured_peer_copy_GBps": 770.0,
                  "formula": "records pushed to other ranks (gx_stats.exchanged_records, exact) * (Kb+2) bytes / exchange_comm (copy-engine "
                             "pushes + arrival barrier, CUDA events on the communication stream; runs under the next chunk's split). "
                             "Hardware NVLink counters read N/A in this pool (profiles/r02q_nvlink_counters.txt)"}

    # ---- e2e through the host API: pinned text in, whole record stream out
    e2e = None
    if not args.no_e2e:
        # the call a user makes: host text in, records out, through a builder that streams its records (serialised slice by
        # slice while the previous slice travels to the host) and takes the text in 64 MiB chunks (H2D of chunk i+1 under
        # the build of chunk i)
        rec_bytes = gb.record_bytes
        ejob = Job(gx, torch, dist, dev, local_rank, rank, world, w, text_np, hint, chunk_bytes=64 << 20, stream_records=True, share=job)
        out_host = torch.empty(max(rec_bytes, 1) + (1 << 20), dtype=torch.uint8).pin_memory()
        ejob.step_host(out_host)
        ejob.barrier()
        t0 = time.perf_counter()
        e_steps = max(1, min(args.steps, 3))
        d2h = 0
        for _ in range(e_steps):
            d2h = ejob.step_host(out_host)
        ejob.barrier()
        dt = time.perf_counter() - t0
        e_phase = ejob.gb.phase_ms()
        ejob.gb.close()
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": tot_occ * e_steps / dt, "unit": "kmers/s", "h2d_bytes_per_step": int(job.host_text.numel()),
               "d2h_bytes_per_step": int(d2h), "steps": e_steps, "timing": "wall clock between synchronises",
               "ms_per_step": dt / e_steps * 1e3, "last_step_phase_ms": e_phase,
               "mode": "64 MiB text chunks double-buffered over PCIe; records streamed (gx_config.reserved[2] bit 0)"}
        del out_host
    job_rounds = job.rounds
    table_info = {"capacity": stats["table_capacity"], "grows": stats["table_grows"], "expected_kmers_hint": hint,
                  "load": stats["distinct_kmers"] / max(1, stats["table_capacity"])}
    job.close()

    # ---- the same build without the capacity hint (the reference CLI has no such option): the table is sized from a pilot
    nohint = None
    if world == 1 and not args.no_nohint:
        j2 = Job(gx, torch, dist, dev, local_rank, rank, world, w, text_np, 0)
        ms2, ph2, _, st2 = j2.timed(max(1, min(args.steps, 3)), 1)
        nohint = {"value": tot_occ / (ms2 * 1e-3), "ms_per_step": ms2, "table": {"capacity": st2["table_capacity"], "grows": st2["table_grows"]},
                  "phase_ms_per_step": ph2, "note": "expected_kmers = 0: first region upserted as a pilot, table sized once from it"}
        j2.close()

    # ---- N > 1: BASELINE configs[3] at full size, strong scaling (the config the >= 5e9 k-mers/s target is quoted on)
    target = None
    if world > 1 and args.target_workload != "none":
        tw = gx.synth.CONFIGS[args.target_workload]
        per_rank = -(-tw.n_reads // world)
        first = rank * per_rank
        mine = max(0, min(per_rank, tw.n_reads - first))
        t_text = gx.synth.shard_text(tw, rank, mine, first_record=first)
        t_occ = gx.synth.occurrences(tw, mine)
        t_hint = expected_distinct(tw, tw.n_reads) // world
        tj = Job(gx, torch, dist, dev, local_rank, rank, world, tw, t_text, t_hint, stream_records=args.target_stream)
        if args.target_stream:
            t_ms, t_phase, t_launch, t_stats = tj.timed_streamed(max(1, args.target_steps), 1)
        else:
            t_ms, t_phase, t_launch, t_stats = tj.timed(max(1, args.target_steps), 1)
        g_occ, g_distinct, g_recbytes = tj.allsum(t_occ, t_stats["distinct_kmers"], t_stats["record_bytes"])
        target = {"workload": workload_desc(tw, 1), "scaling": "strong", "value": g_occ / (t_ms * 1e-3), "unit": "kmers/s",
                  "ms_per_step": t_ms, "steps": max(1, args.target_steps), "warmup": 1, "kmer_occurrences_per_step": g_occ,
                  "distinct_kmers": g_distinct, "record_bytes": g_recbytes, "exchange_rounds_per_step": tj.rounds,
                  "phase_ms_per_step": t_phase, "table": {"capacity": t_stats["table_capacity"], "grows": t_stats["table_grows"]},
                  "target": ">= 5e9 k-mers/s on 8 x B200 at k=55 (BASELINE.json north_star)",
                  "timing": ("wall clock, records streamed slice by slice into a 256 MiB pinned host buffer (the record stream does "
                             "not fit in HBM next to the table)") if args.target_stream else "CUDA events, records resident in HBM"}
        tj.close()

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle
        cores = os.cpu_count() or 1
        n_s = min(n_reads, args.cpu_sample_reads)
        sample_text = gx.synth.readid_text(gx.synth.CONFIGS[args.workload], n_reads=n_s)
        t0 = time.perf_counter()
        _, st = c_oracle.build_graph_records(w.k, sample_text, cores)
        dt = time.perf_counter() - t0
        cpu = {"value": st["occurrences"] / dt, "unit": "kmers/s", "cores": cores, "kind": "port",
               "sample": f"first {n_s} reads of {args.workload} ({st['occurrences']} k-mer occurrences), {dt:.1f} s, "
                         "C port of the reference pipeline (oracle/gx_oracle.c)"}

    if rank == 0:
        line = {
            "metric": "kmer_occurrences_per_sec_graph_build", "value": value, "unit": "kmers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_desc(w, 1), "k": w.k, "read_len": w.read_len, "reads_per_gpu": n_reads,
                       "kmer_occurrences_per_step": tot_occ, "distinct_kmers": tot_distinct,
                       "l2": "inputs (text + hash table) larger than L2; no flush needed",
                       "parallelism": f"hash-partitioned x{world}, one pipelined exchange round per 64 MiB chunk ({job_rounds} per step)" if world > 1 else "single GPU"},
            "bases_per_sec": tot_bases / (ms_per_step * 1e-3),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches * args.steps),
            "roofline": roof, "cpu_baseline": cpu,
            "phase_ms_per_step": phase,
            "table": table_info,
        }
        if nohint is not None:
            line["no_hint"] = nohint
        if nvlink is not None:
            line["nvlink"] = nvlink
        if parity is not None:
            line["parity"] = parity
        if target is not None:
            line["target_workload"] = target
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
