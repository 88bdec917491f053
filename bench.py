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
        "config": {"workload": workload_desc(base, 1), "k": base.k, "read_len": base.read_len},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--ref-sample-reads", type=int, default=60000)
    ap.add_argument("--cpu-sample-reads", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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

    w, text_np, n_reads = make_workload(args.workload, rank, world)
    n_occ = gx.synth.occurrences(w, n_reads)
    n_bases = n_reads * w.read_len * (2 if w.paired else 1)
    host_text = torch.from_numpy(text_np).pin_memory()
    dev_text = host_text.to(dev, non_blocking=False)
    hint = expected_distinct(w, n_reads * world) // world

    gb = gx.GraphBuilder(w.k, device=local_rank, rank=rank, n_ranks=world, expected_kmers=hint)
    stream = torch.cuda.current_stream(dev)
    gb.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(gb.mg_unique_id().copy())
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        gb.mg_init(uid.cpu().numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        gb.reset()
        gb.push_lines_device(dev_text.data_ptr(), dev_text.numel())
        if world > 1:
            gb.mg_exchange()
        gb.finish()

    sampler = ClockSampler(local_rank)
    sampler.start()  # started before the warm-up so that short timed regions still get samples (all under load)
    for _ in range(args.warmup):
        step_device()
    stats = gb.stats()
    barrier()
    launches0 = gb.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase_acc = {}
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
        for key, val in gb.phase_ms().items():
            phase_acc[key] = phase_acc.get(key, 0.0) + val
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = gb.kernel_launches - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        agg = torch.tensor([float(n_occ), float(n_bases), float(stats["distinct_kmers"]), float(launches)], device=dev,
                           dtype=torch.float64)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        tot_occ, tot_bases, tot_distinct, launches = (float(x) for x in agg.tolist())
    else:
        tot_occ, tot_bases, tot_distinct = float(n_occ), float(n_bases), float(stats["distinct_kmers"])
    ms_per_step = ms_total / args.steps
    value = tot_occ / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (extract+insert), timed live with CUDA events inside the library
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    job_bytes, insert_bytes, b_occ, b_dist = algorithmic_bytes(w.k, w.read_len, n_occ, stats["distinct_kmers"])
    insert_ms = (phase_acc.get("insert", 0.0) + phase_acc.get("split", 0.0)) / args.steps
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{args.workload}_extract_insert_bytes_per_launch")
        except Exception:
            traffic = None
    n_insert_launches = max(1, -(-dev_text.numel() // (64 << 20)))
    roof = {
        "bound": "hbm", "kernel": "extract_kernel<KW,EX_UPSERT> (k-mer extract + hash upsert)" if world == 1 else
        "extract_kernel<KW,EX_ROUTE> (k-mer extract + own-key upsert + routing); received keys: insert_records_kernel",
        "achieved": insert_bytes / (insert_ms * 1e-3) / 1e9 if insert_ms > 0 else None,
        "peak": peak, "unit": "GB/s", "peak_source": peak_src,
        "frac": (insert_bytes / (insert_ms * 1e-3) / 1e9 / peak) if insert_ms > 0 else None,
        "traffic": traffic,
        "algorithmic_bytes_per_step": insert_bytes, "kernel_ms_per_step": insert_ms,
        "launches_per_step": n_insert_launches,
        "bytes_per_occurrence": b_occ, "job_bytes_per_step": job_bytes,
        "job_frac": job_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
    }

    # ---- e2e through the host API: pinned text in, whole record stream out
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        rec_bytes = gb.record_bytes
        out_host = torch.empty(max(rec_bytes, 1) + (1 << 20), dtype=torch.uint8).pin_memory()

        def step_host():
            gb.reset()
            gb.push_lines(host_text)
            if world > 1:
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

        step_host()
        barrier()
        t0 = time.perf_counter()
        e_steps = max(1, min(args.steps, 3))
        d2h = 0
        for _ in range(e_steps):
            d2h = step_host()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": tot_occ * e_steps / dt, "unit": "kmers/s", "h2d_bytes_per_step": int(host_text.numel()),
               "d2h_bytes_per_step": int(d2h), "steps": e_steps, "timing": "wall clock between synchronises"}

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
                       "parallelism": f"hash-partitioned x{world}" if world > 1 else "single GPU"},
            "bases_per_sec": tot_bases / (ms_per_step * 1e-3),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu,
            "phase_ms_per_step": {k_: v / args.steps for k_, v in phase_acc.items()},
            "table": {"capacity": stats["table_capacity"], "grows": stats["table_grows"]},
        }
        print(json.dumps(line), flush=True)
    gb.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
