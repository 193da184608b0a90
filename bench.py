#!/usr/bin/env python
"""bench.py — EM iterations/s of the EMSAR quantification hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference arm: CPU EM (oracle port) on the host cores

Workload (BASELINE.json configs[1]): synthetic human-scale SE rsh index — 200K transcripts, ~2M rsh classes,
30M reads, -k 100 — one sample per GPU (samples are independent: -M sharding, no collective, weak scaling).
A step = `--em-iters` EM iterations over the resident packed sample. `value` = EM iterations/s summed over
ranks; `e2e` = the same metric through the C ABI with HOST buffers (pinned read lists H2D, counting, model
build, the same number of EM iterations, results D2H) inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V2 = dict(T=200000, n_multi=1800000, alpha=2.4, kmax=99, module_cap=5000, p_cross=0.1, scatter=True)
WORKLOADS = {
    # name: generator ("v2" = SURVEY.md section 8d spec at full size, emsar_b200.synth.make_index_v2 / make_reads_fast; "v1" = the
    # family-block generator of the small parity cases), index kwargs, reads per sample
    # BASELINE.json configs[1]: 200K transcripts, 2M classes (mean cardinality 4.47), 10 % of the classes across a paralog family whose
    # gene families lie scattered over the tid range, 30M reads
    "config2_human_se": ("v2", V2, 30_000_000),
    "config2_shuffled": ("v2", dict(V2, shuffle_tids=True), 30_000_000),       # the same transcriptome under random transcript names
    "config2_100m": ("v2", V2, 100_000_000),                                   # north star: 200K / 2M / 100M reads on one B200
    # BASELINE.json configs[2]: PE L101 F101-500 (nF = 400: 3.2 GB of EUMA), 100M reads
    "config3_pe_100m": ("v2", dict(V2, nF=400, frag_min=101, readlength=101), 100_000_000),
    # BASELINE.json configs[4]: -k 1000, cardinality ~ k^-1.5 on [2, 999] (mean 39, nnz 70M), 200 hub transcripts in ~10^4 classes each
    "config5_full": ("v2", dict(V2, alpha=1.5, kmax=999, hubs=200, hub_classes=10000), 10_000_000),
    # round-1 workloads (family blocks of consecutive tids only; kept for continuity with profiles/r1*)
    "config2_r1": ("v1", dict(T=200000, n_multi=2100000, alpha=2.4, kmax=99, module_cap=5000), 30_000_000),
    "config5_stress": ("v1", dict(T=60000, n_multi=250000, alpha=1.5, kmax=999, module_cap=3000, hubs=20, hub_classes=8000), 3_000_000),
    "small": ("v1", dict(T=20000, n_multi=200000, alpha=2.4, kmax=99, module_cap=500), 3_000_000),
    "tiny": ("v1", dict(T=2000, n_multi=20000, alpha=2.4, kmax=40, module_cap=200), 200_000),
}
INDEX_SEED = {"config3_pe_100m": 3, "config5_full": 5}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload, em_iters):
    """DRAM bytes of one k_em_persistent launch from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if t.get("workload") != workload:
            return None
        return int(t["dram_bytes_per_launch"] + t.get("dram_bytes_per_extra_iteration", 0) * em_iters)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.rows = gpu, threading.Event(), []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 6 and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


_INDEX_CACHE = {}


def make_index(name):
    """One index per workload, shared by every sample / rank (cached inside the process)."""
    from emsar_b200 import synth
    if name not in _INDEX_CACHE:
        gen, kw, _ = WORKLOADS[name]
        seed = INDEX_SEED.get(name, 2)
        _INDEX_CACHE.clear()                # one big index at a time
        _INDEX_CACHE[name] = synth.make_index_v2(seed=seed, **kw) if gen == "v2" else synth.make_index(seed=seed, **kw)
    return _INDEX_CACHE[name]


def make_workload(name, seed, n_reads=None):
    from emsar_b200 import synth
    gen, kw, n_default = WORKLOADS[name]
    t0 = time.time()
    idx = make_index(name)
    n = n_default if n_reads is None else n_reads
    reads = synth.make_reads_fast(idx, n, seed=seed) if gen == "v2" else synth.make_reads(idx, n, seed=seed)    # this rank's sample
    return idx, reads, time.time() - t0


def reference_binary_twin(cores):
    """the UNMODIFIED reference binary (oracle/_ref/emsar, compiled from /root/reference by oracle/Makefile) on the
    scaled twin of config #2 (T = 20K, C ~ 190K, 3M reads, modules <= 500): its estimator is quadratic in module size and cannot
    finish the full config. One round (-n 1) instead of its default four. Returns None when the binary did not travel here."""
    import re
    import tempfile
    from emsar_b200 import synth
    ref = os.path.join(ROOT, "oracle", "_ref", "emsar")
    if not os.path.exists(ref):
        return None
    idx, reads, _ = make_workload("small", seed=1000)
    d = tempfile.mkdtemp(prefix="emsar_twin_")
    synth.write_rsh(idx, d + "/x.rsh")
    synth.write_bowtie_se(idx, reads, d + "/x.bowtie")
    t0 = time.perf_counter()
    out = subprocess.run([ref, "-p", str(cores), "-n", "1", "-I", d + "/x.rsh", d + "/out", "p", d + "/x.bowtie"], capture_output=True, text=True)
    sec = time.perf_counter() - t0
    import shutil
    shutil.rmtree(d, ignore_errors=True)
    hms = [int(h) * 3600 + int(m) * 60 + int(s_) for h, m, s_ in re.findall(r"\d\d/\d\d,(\d\d):(\d\d):(\d\d)", out.stdout)]
    mle = None
    m = re.search(r"round 1/1\.\.\.\n\d\d/\d\d,(\d\d):(\d\d):(\d\d)(?s:.*?)computing effective length[^\n]*\n\d\d/\d\d,(\d\d):(\d\d):(\d\d)", out.stdout)
    if m:
        a = [int(x) for x in m.groups()]
        mle = (a[3] * 3600 + a[4] * 60 + a[5]) - (a[0] * 3600 + a[1] * 60 + a[2])
    return {"seconds": sec, "samples_per_min": 60.0 / sec, "mle_seconds": mle, "rc": out.returncode, "threads": cores,
            "workload": f"scaled twin: T={idx.T} C={idx.C} reads={len(reads.read_fraglen)}, modules <= 500 transcripts, -n 1"}


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path: no EM exists in parklab/emsar, so the
    metric 'EM iterations/s' is timed on the oracle port (same update as the CUDA kernel) with all host threads;
    oracle/_ref/emsar (the real binary) is timed on a scaled twin for the samples/min context figure."""
    if rank != 0:
        return
    from oracle import oracle
    idx, reads, gen_s = make_workload(args.workload, seed=1000)
    cores = os.cpu_count() or 1
    R, F, N = oracle.count(idx, reads)
    Wf, adj, ps, iE = oracle.prepare(idx, F, N)
    one = oracle.em_time(idx, R, ps, 2, cores) / 2
    iters = max(2, min(args.em_iters, int(4.0 / max(one, 1e-6))))       # bounded sample: ~4 s of CPU work per step
    for _ in range(args.warmup):
        oracle.em_time(idx, R, ps, 1, cores)
    t = 0.0
    for _ in range(args.steps):
        t += oracle.em_time(idx, R, ps, iters, cores)
    v = args.steps * iters / t
    line = {"impl": "reference", "metric": "em_iterations_per_sec", "value": v, "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "T": idx.T, "C": idx.C, "reads": int(N), "em_iters_per_step": iters},
            "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": cores, "kind": "port",
                             "sample": f"{iters} EM iterations per step of the full {args.workload} model, pthread team of {cores}"},
            "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_ref_binary:
        try:
            line["reference_binary"] = reference_binary_twin(cores)      # `emsar -p N` itself, timed in the same run
        except Exception as e:                                           # a reported side figure: never lose the line over it
            line["reference_binary"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2_human_se", choices=list(WORKLOADS))
    ap.add_argument("--em-iters", type=int, default=2000, help="EM iterations per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-converge", action="store_true", help="skip the one-off run to convergence (samples/min)")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the host-buffer leg (e2e is then null)")
    ap.add_argument("--no-ref-binary", action="store_true",
                    help="--impl reference only: skip timing the unmodified reference binary (oracle/_ref/emsar -p N) on the scaled twin (about 1-2 minutes)")
    ap.add_argument("--shard", default="samples", choices=["samples", "classes"],
                    help="N>1: independent samples per GPU (-M list, weak scaling, the default) or ONE sample whose classes are "
                         "range-sharded over the GPUs with the per-iteration all-reduce (BASELINE.json configs[2], strong scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from emsar_b200.api import Context, Index

    by_class = args.shard == "classes" and world > 1
    idx, reads, gen_s = make_workload(args.workload, seed=1000 if by_class else 1000 + rank)
    ctx = Context(local)
    if by_class:
        ctx.comm_init_torch()
        n_all = len(reads.read_fraglen)                      # every rank counts its slice of the read groups
        lo, hi = n_all * rank // world, n_all * (rank + 1) // world
        base = int(reads.read_ptr[lo])
        reads.read_tid = reads.read_tid[base:int(reads.read_ptr[hi])]
        reads.read_ptr = reads.read_ptr[lo:hi + 1] - base
        reads.read_fraglen = reads.read_fraglen[lo:hi]
    ix = Index(ctx, idx)
    # pinned host copies of this rank's read lists (the e2e leg copies them every step)
    h_ptr = torch.from_numpy(reads.read_ptr).pin_memory()
    h_tid = torch.from_numpy(reads.read_tid).pin_memory()
    h_fl = torch.from_numpy(reads.read_fraglen).pin_memory()
    h2d_bytes = h_ptr.numel() * 8 + h_tid.numel() * 4 + h_fl.numel() * 4
    # resident sample for the device-timed leg
    def load(s_):
        s_.count(h_ptr, h_tid, h_fl)
        if by_class:
            s_.counts_allreduce()
        s_.prepare(sharded=by_class)

    smp = ix.sample()
    load(smp)
    st = smp.model_stats()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def step_resident():
        flush.zero_()                      # L2 flush between steps (torch stream), then the EM loop (library stream)
        torch.cuda.synchronize()
        it, fd, ms = smp.em_run(max_iter=args.em_iters, stop_on_conv=False, reset_theta=True)
        return it, ms

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = ctx.launches()
    t0 = time.perf_counter()
    iters_done, em_ms = 0, 0.0
    for _ in range(args.steps):
        it, ms = step_resident()
        iters_done += it
        em_ms += ms
    barrier()
    t1 = time.perf_counter()
    launches = ctx.launches() - l0
    wall = t1 - t0
    # device time of the EM kernel alone (CUDA events on the library's stream) -> roofline
    t_em = em_ms / 1e3

    # ---- e2e leg: host buffers through the C ABI ----
    def step_e2e():
        s = ix.sample()
        load(s)
        it, fd, ms = s.em_run(max_iter=args.em_iters, stop_on_conv=False)
        r = s.finalize()
        s.close()
        return it, r

    e_iters, e_wall = 0, 1.0
    if not args.no_e2e:
        step_e2e()
        barrier()
        e0 = time.perf_counter()
        for _ in range(args.steps):
            it, r = step_e2e()
            e_iters += it
        barrier()
        e1 = time.perf_counter()
        e_wall = e1 - e0
    d2h_bytes = idx.T * (8 * 4 + 4)

    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    # ---- one sample to convergence (samples/min context) ----
    conv = None
    if not args.no_converge:
        barrier()
        c0 = time.perf_counter()
        s = ix.sample()
        s.count(h_ptr, h_tid, h_fl)
        if by_class:
            s.counts_allreduce()
        r = s.solve(sharded=by_class)
        s.close()
        c1 = time.perf_counter()
        conv_s = c1 - c0
        if world > 1:                                       # every rank solves a sample (or the ranks share one): the slowest decides
            tc = torch.tensor([conv_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            conv_s = float(tc.item())
        conv = {"seconds": conv_s, "samples_per_min": (1 if by_class else world) * 60.0 / conv_s,
                "what": "host read lists -> counts -> model -> EM to convergence -> FPKM/TPM on the host, per sample", "n_iter": int(r["n_iter"]), "final_delta": float(r["final_delta"]), "em_ms": float(r["em_ms"]),
                "prep_ms": float(r["prep_ms"])}

    # ---- reduce over ranks: max time, summed work ----
    if world > 1:
        tt = torch.tensor([wall, e_wall, t_em], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall, e_wall, t_em = tt.tolist()
        ww = torch.tensor([iters_done, e_iters, launches, st["bytes_per_iter"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        iters_tot, e_iters_tot, launches_tot, bytes_all = ww.tolist()
        if by_class:                       # one sample: the ranks iterate together
            iters_tot, e_iters_tot = iters_done, e_iters
    else:
        iters_tot, e_iters_tot, launches_tot, bytes_all = iters_done, e_iters, launches, st["bytes_per_iter"]

    if rank == 0:
        peak, peak_src = peaks()
        achieved = st["bytes_per_iter"] * iters_done / t_em / 1e9 if t_em > 0 else 0.0
        if by_class:                       # the whole sample's bytes per iteration against the N GPUs' bandwidth
            achieved = bytes_all * iters_done / t_em / 1e9 if t_em > 0 else 0.0
            peak *= world
            st = dict(st, bytes_per_iter=int(bytes_all), rank0_bytes_per_iter=st["bytes_per_iter"])
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import oracle
            cores = os.cpu_count() or 1
            R, F, N = smp.counts()
            Wf, adj, ps, iE = oracle.prepare(idx, F, N)
            one = oracle.em_time(idx, R, ps, 2, cores) / 2
            n_it = max(2, int(10.0 / max(one, 1e-6)))
            n_it = min(n_it, 2000)
            sec = oracle.em_time(idx, R, ps, n_it, cores)
            cpu = {"value": n_it / sec, "unit": "iterations/s", "cores": cores, "kind": "port",
                   "sample": f"{n_it} EM iterations of the same packed sample ({sec:.1f} s), pthread team of {cores}"}
        line = {
            "metric": "em_iterations_per_sec", "value": iters_tot / wall, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "strong" if by_class else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "T": idx.T, "C": idx.C, "reads_per_sample": int(len(reads.read_fraglen)) * (world if by_class else 1),
                       "C_a": st["C_a"], "nnz_a": st["nnz_a"], "em_iters_per_step": args.em_iters, "samples": 1 if by_class else world,
                       "parallelism": (f"one sample, classes range-sharded x{world}, per-iteration fp64 all-reduce over NVLink "
                                       f"({'inside the EM kernel, peer memory' if ctx.comm_info()['peer_memory'] == 1 else 'ncclAllReduce'})")
                       if by_class else f"sample-sharded x{world} (-M), no collective",
                       "l2": "flushed between steps (256 MiB memset); iterations inside a step reuse L2 as the production loop does"},
            "e2e": None if args.no_e2e else {"value": e_iters_tot / e_wall, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": 1e3 * e_wall / args.steps},
            "gpu_launches": int(launches_tot),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if world > 1 else measured_traffic(args.workload, args.em_iters),
                         "traffic_note": "DRAM bytes per launch (ncu, profiles/traffic.json); algorithmic bytes per launch = bytes_per_iter x "
                                         "em_iters_per_step: the packed model is L2-resident after the first iteration",
                         "algorithmic_bytes_per_launch": int(st["bytes_per_iter"]) * int(args.em_iters),
                         "peak_source": peak_src, "kernel": "k_em_persistent", "bytes_per_iter": st["bytes_per_iter"],
                         "us_per_iter": 1e6 * t_em / max(iters_done, 1)},
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "to_convergence": conv,
            "model": st, "gen_seconds": gen_s,
        }
        print(json.dumps(line), flush=True)
    smp.close()
    ix.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
