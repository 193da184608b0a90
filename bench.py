#!/usr/bin/env python
"""bench.py — EM iterations/s of the EMSAR quantification hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference arm: CPU EM (oracle port) on the host cores

Workload of the headline line (BASELINE.json configs[1], generator of SURVEY.md section 8d): synthetic human-scale SE rsh index —
200K transcripts, 2M rsh classes (10 % of them across a paralog family scattered over the tid range), 30M reads, -k 100 — one
sample per GPU (samples are independent: -M sharding, no collective, weak scaling).
A step = `--em-iters` EM iterations over the resident packed sample. `value` = EM iterations/s summed over ranks; `e2e` = the same
metric through the C ABI with HOST buffers (pinned read lists H2D, counting, model build, the same number of EM iterations,
results D2H) inside the timed region.

Further objects of the JSON line (all measured in the same run, none of them inside the timed region of `value`):
  kernels          roofline entries of the one-off kernels (k_count; k_adjeuma_stream comes with config3 below)
  other_workloads  N = 1 only: the other BASELINE configs at their stated size (config3 PE 100M reads, config5 -k 1000 full size,
                   the 100M-read north-star sample, config2 under random transcript names): it/s, roofline fraction, kernel variant
  m64              BASELINE configs[3]: a -M batch of 8 samples per GPU (64 on 8 GPUs) sharing one index, largest-first assignment,
                   host buffers -> results: samples/min
  class_sharded    N > 1 only: ONE sample sharded over the N GPUs (BASELINE configs[2]) with its parity against the single-GPU solve. Each
                   workload runs in child processes (one per GPU, their own process group) under a time limit, a small case first: a
                   stall of the cross-GPU kernel at some GPU count shows up as {"error": ...} here instead of a run without a line
  file_to_file     N = 1 only: this repo's `emsar` command and the unmodified reference binary on the same .rsh + bowtie text files
  index_build      N = 1 only: the rsh index of a generated transcriptome constructed on the device, by the host builder and by the
                   reference's emsar-build, files compared
Progress goes to stderr per rank ("[bench r<rank> +<seconds>s] ..."). A run that has not printed its line after EMSAR_BENCH_LIMIT_S
(default 1200) seconds prints the headline it has, marked "incomplete", and ends.
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
OTHER_WORKLOADS = ["config2_100m", "config2_shuffled", "config3_pe_100m", "config5_full"]
L2_NOTE = "flushed between steps (256 MiB memset); iterations inside a step reuse L2 as the production loop does"


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


def make_reads(name, idx, seed, n_reads=None):
    from emsar_b200 import synth
    gen, _, n_default = WORKLOADS[name]
    n = n_default if n_reads is None else n_reads
    return synth.make_reads_fast(idx, n, seed=seed) if gen == "v2" else synth.make_reads(idx, n, seed=seed)


def make_workload(name, seed, n_reads=None):
    t0 = time.time()
    idx = make_index(name)
    reads = make_reads(name, idx, seed, n_reads)         # this rank's sample
    return idx, reads, time.time() - t0


def config_dict(workload, idx, reads_per_sample, C_a, nnz_a, em_iters, world):
    """The `config` object: the same keys and values in the repo arm and in the reference arm."""
    return {"workload": workload, "T": int(idx.T), "C": int(idx.C), "reads_per_sample": int(reads_per_sample), "C_a": int(C_a), "nnz_a": int(nnz_a),
            "em_iters_per_step": int(em_iters), "samples": int(world), "parallelism": f"sample-sharded x{world} (-M), no collective", "l2": L2_NOTE}


def active_model_size(idx, R, ps):
    """C_a, nnz_a of the packed sample (multi-tid classes with reads that are in the model)."""
    T = idx.T
    act = (R[T:] > 0) & (ps[T:] > 0)
    k = np.diff(idx.class_ptr)[T:]
    return int(act.sum()), int(k[act].sum())


# ---------------------------------------------------------------------------------------------------------------------
# reference side
# ---------------------------------------------------------------------------------------------------------------------
def write_twin(d):
    """The scaled twin of config #2 (T = 20K, C ~ 190K, 3M reads, modules <= 500) as the files a user has: .rsh + bowtie text."""
    from emsar_b200 import synth
    idx, reads, _ = make_workload("small", seed=1000)
    synth.write_rsh(idx, d + "/x.rsh")
    synth.write_bowtie_se(idx, reads, d + "/x.bowtie")
    return idx, reads


def run_reference_binary(d, cores):
    """oracle/_ref/emsar (the UNMODIFIED reference, compiled from /root/reference by oracle/Makefile) on the twin's files. Its estimator
    is quadratic in module size and cannot finish the full config; one round (-n 1) instead of its default four."""
    import re
    ref = os.path.join(ROOT, "oracle", "_ref", "emsar")
    if not os.path.exists(ref):
        return None
    t0 = time.perf_counter()
    out = subprocess.run([ref, "-p", str(cores), "-n", "1", "-I", d + "/x.rsh", d + "/ref_out", "p", d + "/x.bowtie"], capture_output=True, text=True)
    sec = time.perf_counter() - t0
    mle = None
    m = re.search(r"round 1/1\.\.\.\n\d\d/\d\d,(\d\d):(\d\d):(\d\d)(?s:.*?)computing effective length[^\n]*\n\d\d/\d\d,(\d\d):(\d\d):(\d\d)", out.stdout)
    if m:
        a = [int(x) for x in m.groups()]
        mle = (a[3] * 3600 + a[4] * 60 + a[5]) - (a[0] * 3600 + a[1] * 60 + a[2])
    return {"seconds": sec, "samples_per_min": 60.0 / sec, "mle_seconds": mle, "rc": out.returncode, "threads": cores}


def read_fpkm(path):
    rows = [l.rstrip("\n").split("\t") for l in open(path)][1:]
    return np.array([float(r[1]) for r in rows]), np.array([float(r[4]) for r in rows])


def reference_binary_twin(cores):
    import shutil
    import tempfile
    d = tempfile.mkdtemp(prefix="emsar_twin_")
    try:
        idx, reads = write_twin(d)
        r = run_reference_binary(d, cores)
        if r is not None:
            r["workload"] = f"scaled twin: T={idx.T} C={idx.C} reads={len(reads.read_fraglen)}, modules <= 500 transcripts, -n 1"
        return r
    finally:
        shutil.rmtree(d, ignore_errors=True)


def file_to_file_twin(cores, device):
    """What a user sees: `emsar -I x.rsh out p x.bowtie` of this repo against the unmodified reference binary on the SAME files, wall
    clock, parsing and output files included."""
    import shutil
    import tempfile
    d = tempfile.mkdtemp(prefix="emsar_f2f_")
    try:
        idx, reads = write_twin(d)
        ours = os.path.join(ROOT, "emsar_b200", "bin", "emsar")
        env = dict(os.environ, EMSAR_DEVICES=str(device))
        subprocess.run([ours, "-q", "-I", d + "/x.rsh", d + "/warm", "p", d + "/x.bowtie"], capture_output=True, text=True, env=env)      # CUDA context + page cache
        t0 = time.perf_counter()
        o = subprocess.run([ours, "-q", "-I", d + "/x.rsh", d + "/our_out", "p", d + "/x.bowtie"], capture_output=True, text=True, env=env)
        ours_s = time.perf_counter() - t0
        res = {"workload": f"scaled twin of config #2 as files: T={idx.T} C={idx.C} reads={len(reads.read_fraglen)}", "ours_seconds": ours_s, "ours_rc": o.returncode,
               "ours_samples_per_min": 60.0 / ours_s}
        ref = run_reference_binary(d, cores)
        if ref is not None and ref["rc"] == 0 and o.returncode == 0:
            fo, ro = read_fpkm(d + "/our_out/p.0.fpkm")
            fr, rr = read_fpkm(d + "/ref_out/p.0.fpkm")
            ok = (np.abs(fo - fr) <= 1e-6 * np.abs(fr) + 2e-6) | (np.abs(ro - rr) <= 1e-3)
            res.update({"reference_seconds": ref["seconds"], "reference_threads": ref["threads"], "reference_rounds": 1, "speedup": ref["seconds"] / ours_s,
                        "fpkm_within_tolerance": float(ok.mean()), "max_abs_ireadcount_diff": float(np.abs(ro - rr).max()),
                        "note": "reference = pattern search from a random start (one round), not reproducible run to run (SURVEY.md 0.2); tolerance max(1e-6 rel, 1e-3 reads) at the files' 6 decimals"})
        elif ref is None:
            res["reference_seconds"] = None
        return res
    finally:
        shutil.rmtree(d, ignore_errors=True)


def index_build_leg(cores, device, transcripts=20000):
    """the rsh index of a generated transcriptome (gene families over a shared exon pool, 1.7K bases per transcript) built three
    ways - classes constructed on the device (`emsar-build --device`), by the host builder, and by the unmodified reference `emsar-build`
    (oracle/_ref, on a quarter of the transcripts: it is the slow one) - with the files compared byte for byte (SHA-256)."""
    import hashlib
    import shutil
    import tempfile
    from profiles import build_bench as bb
    mine = os.path.join(ROOT, "emsar_b200", "bin", "emsar-build")
    ref = os.path.join(ROOT, "oracle", "_ref", "emsar-build")
    d = tempfile.mkdtemp(prefix="emsar_build_")

    def run(tool, extra, fa, tag, limit):
        t0 = time.perf_counter()
        try:
            r = subprocess.run([tool, "-q"] + extra + [fa, "50", os.path.join(d, tag), "x"], capture_output=True, text=True, timeout=limit,
                               env=dict(os.environ, EMSAR_BUILD_TIMING="1"))
        except subprocess.TimeoutExpired:
            return {"error": f"over {limit} s"}
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": (r.stdout + r.stderr)[-300:]}
        data = open(os.path.join(d, tag, "x.rsh"), "rb").read()
        return {"seconds": dt, "sha256": hashlib.sha256(data).hexdigest()[:16], "classes": data.count(b"\n"),
                "stages": [l.split(": ", 1)[1] for l in r.stderr.splitlines() if l.startswith("build timing")]}

    try:
        fa, fa4 = os.path.join(d, "t.fa"), os.path.join(d, "t4.fa")
        bases = bb.make_fasta(fa, transcripts)
        bases4 = bb.make_fasta(fa4, max(transcripts // 4, 1))
        out = {"workload": f"single-end unstranded, read length 50, {transcripts} transcripts / {bases} bases (reference: {max(transcripts // 4, 1)} / {bases4})"}
        run(mine, ["--device", str(device)], fa4, "warm", 300)                     # CUDA start-up, page cache
        out["device"] = run(mine, ["--device", str(device)], fa, "dev", 600)
        out["host"] = run(mine, ["-p", str(cores)], fa, "host", 600)
        out["device_quarter"] = run(mine, ["--device", str(device)], fa4, "dev4", 300)
        out["reference_quarter"] = run(ref, [], fa4, "ref4", 120) if os.path.exists(ref) else None
        out["identical_device_host"] = out["device"].get("sha256") is not None and out["device"].get("sha256") == out["host"].get("sha256")
        if out["reference_quarter"] is not None:
            out["identical_device_reference"] = out["device_quarter"].get("sha256") is not None and \
                out["device_quarter"].get("sha256") == out["reference_quarter"].get("sha256")
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


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
    C_a, nnz_a = active_model_size(idx, R, ps)
    iters = args.em_iters                                 # same step as the repo arm (about 4 s of CPU work per step on 16 cores)
    for _ in range(args.warmup):
        oracle.em_time(idx, R, ps, max(1, iters // 100), cores)
    t = 0.0
    for _ in range(args.steps):
        t += oracle.em_time(idx, R, ps, iters, cores)
    v = args.steps * iters / t
    line = {"impl": "reference", "metric": "em_iterations_per_sec", "value": v, "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, idx, len(reads.read_fraglen), C_a, nnz_a, iters, world),
            "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": cores, "kind": "port",
                             "sample": f"{iters} EM iterations per step of the full {args.workload} model, pthread team of {cores} "
                                       f"(warm-up steps run {max(1, iters // 100)} iterations each)"},
            "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_ref_binary:
        try:
            line["reference_binary"] = reference_binary_twin(cores)      # `emsar -p N` itself, timed in the same run
        except Exception as e:                                           # a reported side figure: never lose the line over it
            line["reference_binary"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# legs of the repo arm
# ---------------------------------------------------------------------------------------------------------------------
def count_roofline(ctx, smp_factory, reads, torch):
    """k_count with its inputs resident in HBM, timed with CUDA events on the library's stream. Algorithmic bytes: the
    read lists (12 + 4 k_r per read) plus, per read, one 32-byte hash / singleton sector, the key verification (4 k_r for k_r > 1)
    and one 32-byte counter sector."""
    d_ptr = torch.from_numpy(reads.read_ptr).cuda()
    d_tid = torch.from_numpy(reads.read_tid).cuda()
    d_fl = torch.from_numpy(reads.read_fraglen).cuda()
    torch.cuda.synchronize()
    k = np.diff(reads.read_ptr)
    n, tot = len(k), int(k.sum())
    alg = 12 * n + 4 * tot + 64 * n + 4 * int(k[k > 1].sum())
    best = None
    for rep in range(3):
        s = smp_factory()
        ctx.synchronize()
        ctx.timer_start()
        s.count(d_ptr, d_tid, d_fl)
        ms = ctx.timer_stop()
        s.close()
        best = ms if best is None else min(best, ms)
    peak, src = peaks()
    del d_ptr, d_tid, d_fl
    torch.cuda.empty_cache()
    return {"kernel": "k_count", "bound": "hbm", "ms": best, "reads": n, "reads_per_sec": n / (best * 1e-3), "algorithmic_bytes": alg,
            "achieved": alg / (best * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (best * 1e-3) / 1e9 / peak, "peak_source": src,
            "note": "random 32-byte sectors (hash probe, key verify, counter): a sector rate, not a stream"}


def em_leg(ctx, s, iters, steps, warmup, torch, flush):
    for _ in range(warmup):
        s.em_run(max_iter=iters, stop_on_conv=False, reset_theta=True)
    n_it, ms_tot = 0, 0.0
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        it, fd, ms = s.em_run(max_iter=iters, stop_on_conv=False, reset_theta=True)
        n_it += it
        ms_tot += ms
    return n_it, ms_tot


def other_workload(ctx, name, torch, flush, em_iters):
    """One of the other BASELINE configs at its stated size: counting, model build and the EM kernel, device-timed."""
    from emsar_b200.api import Index
    t0 = time.time()
    idx, reads, gen_s = make_workload(name, seed=1000)
    ix = Index(ctx, idx)
    s = ix.sample()
    s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
    s.prepare()
    st = s.model_stats()
    peak, src = peaks()
    out = {"workload": name, "T": int(idx.T), "C": int(idx.C), "nF": int(idx.nF), "reads": int(len(reads.read_fraglen)), "read_tids": int(len(reads.read_tid)),
           "C_a": st["C_a"], "nnz_a": st["nnz_a"], "em_variant": st["em_variant"], "all_local": st["all_local"], "bytes_per_iter": st["bytes_per_iter"],
           "index_bytes": st["index_bytes"]}
    if idx.nF > 1:
        ms = s.time_adjeuma(3)
        alg = 4 * idx.C * idx.nF
        out["k_adjeuma"] = {"kernel": "k_adjeuma_stream", "bound": "hbm", "ms": ms, "algorithmic_bytes": int(alg), "achieved": alg / (ms * 1e-3) / 1e9,
                            "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak}
    iters = em_iters if st["bytes_per_iter"] < 3e8 else max(50, em_iters // 10)
    n_it, ms_tot = em_leg(ctx, s, iters, 2, 1, torch, flush)
    ach = st["bytes_per_iter"] * n_it / (ms_tot * 1e-3) / 1e9
    out.update({"value": n_it / (ms_tot * 1e-3), "unit": "iterations/s", "us_per_iter": 1e3 * ms_tot / n_it,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src}})
    it, fd, ms = s.em_run(stop_on_conv=True)              # from where the timed steps left off to convergence: the stopping rule holds at this size
    r = s.finalize()
    out["to_convergence"] = {"n_iter": int(r["n_iter"]), "final_delta": float(fd), "tpm_sum": float(r["tpm"].sum())}
    s.close()
    ix.close()
    out["seconds"] = time.time() - t0
    return out


def m64_leg(ctx, ix, idx, workload, rank, world, torch, dist, per_gpu):
    """BASELINE configs[3]: a -M batch of `per_gpu` x N samples sharing one index (64 on 8 GPUs), read counts U[20M, 40M], assigned
    largest first to the least loaded GPU; every sample goes host read lists -> counts -> model -> EM to convergence -> results."""
    n_samples = per_gpu * world
    rng = np.random.default_rng(4)
    sizes = rng.integers(20_000_000, 40_000_001, size=n_samples)
    load, mine = np.zeros(world), []
    for j in np.argsort(-sizes, kind="stable"):                  # LPT: the same list on every rank
        r = int(np.argmin(load))
        load[r] += sizes[j]
        if r == rank:
            mine.append(int(j))
    # whatever happens on one rank, every rank reaches the collectives below (a rank that skipped them would leave the others waiting)
    bufs, err = [], None
    try:
        for j in mine:
            rd = make_reads(workload, idx, seed=1000 + j, n_reads=int(sizes[j]))
            bufs.append((j, torch.from_numpy(np.diff(rd.read_ptr).astype(np.uint16)).pin_memory(), torch.from_numpy(rd.read_tid).pin_memory(),
                         None if idx.nF == 1 else torch.from_numpy(rd.read_fraglen.astype(np.uint16)).pin_memory(), int(rd.read_fraglen[0])))
            del rd
    except Exception as e:
        err = repr(e)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    iters, h2d, chk, done, not_conv = 0, 0, 0.0, 0, 0
    try:
        for j, p, t, f, f0 in bufs:
            s = ix.sample()
            s.count_compact(p, t, f, f0)                     # the compact wire form: uint16 lengths, tids, uint16 / no fragment lengths
            r = s.solve()
            s.close()
            iters += int(r["n_iter"])
            h2d += p.numel() * 2 + t.numel() * 4 + (0 if f is None else f.numel() * 2)
            chk += float(r["tpm"].sum())
            done += 1
            not_conv += 0 if r["final_delta"] <= 1.0 else 1
        ctx.synchronize()
    except Exception as e:
        err = err or repr(e)
    mine_s = time.perf_counter() - t0
    if err is not None:
        note(f"m64 leg failed on this rank: {err}")
    tt = torch.tensor([mine_s], dtype=torch.float64, device="cuda")
    ss = torch.tensor([float(iters), float(h2d), float(done), float(not_conv), 0.0 if err is None else 1.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ss)
    sec = float(tt.item())
    iters_all, h2d_all, n_all, nc_all, n_err = ss.tolist()
    if n_err > 0:
        return {"workload": workload, "samples": int(n_all), "error": err or "a rank failed (its stderr has the exception)"}
    return {"workload": f"{workload}: {n_samples} samples (seeds 1000..{999 + n_samples}) sharing one index, 20-40M reads each", "samples": int(n_all), "samples_per_gpu": per_gpu,
            "assignment": "largest first to the least loaded GPU (LPT)", "seconds": sec, "samples_per_min": 60.0 * n_all / sec, "em_iterations": int(iters_all),
            "h2d_bytes": int(h2d_all), "not_converged": int(nc_all),
            "what": "pinned host read lists -> counts -> model -> EM to convergence -> FPKM/TPM on the host; max over ranks"}


def bcast_workload(name, seed, rank, torch, dist):
    """Rank 0 generates the index and the sample, the other ranks receive them over NCCL (generation is host work; eight copies of it would
    only fight over the host cores)."""
    import types
    from emsar_b200 import synth
    if rank == 0:
        idx, reads, _ = make_workload(name, seed)
        meta = [dict(T=idx.T, min_fraglength=idx.min_fraglength, max_fraglength=idx.max_fraglength, readlength=idx.readlength, max_t_size=idx.max_t_size,
                     shapes=[a.shape for a in (idx.class_ptr, idx.class_tid, idx.euma, idx.has_node, reads.read_ptr, reads.read_tid, reads.read_fraglen)])]
    else:
        meta = [None]
    dist.broadcast_object_list(meta, 0)
    m = meta[0]
    dts = [torch.int64, torch.int32, torch.int32, torch.uint8, torch.int64, torch.int32, torch.int32]
    arrs = []
    src = (idx.class_ptr, idx.class_tid, idx.euma, idx.has_node, reads.read_ptr, reads.read_tid, reads.read_fraglen) if rank == 0 else [None] * 7
    for a, shp, dt in zip(src, m["shapes"], dts):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda() if rank == 0 else torch.empty(tuple(shp), dtype=dt, device="cuda")
        dist.broadcast(t, 0)
        arrs.append(t.cpu().numpy())
        del t
    torch.cuda.empty_cache()
    idx2 = synth.SynthIndex(T=m["T"], names=None, class_ptr=arrs[0], class_tid=arrs[1], euma=arrs[2], has_node=arrs[3], min_fraglength=m["min_fraglength"],
                            max_fraglength=m["max_fraglength"], readlength=m["readlength"], max_t_size=m["max_t_size"])
    return idx2, types.SimpleNamespace(read_ptr=arrs[4], read_tid=arrs[5], read_fraglen=arrs[6])


def class_sharded_leg(ctx, name, idx, reads, rank, world, torch, dist):
    """ONE sample sharded over the N GPUs (BASELINE configs[2]): every rank counts its slice of the read groups, the integer counts are
    summed, the EM runs to convergence sharded; rank 0 then solves the same sample alone and the two answers are compared."""
    from emsar_b200.api import Index
    ix = Index(ctx, idx)
    n = len(reads.read_fraglen)
    lo, hi = n * rank // world, n * (rank + 1) // world
    base = int(reads.read_ptr[lo])
    s = ix.sample()
    s.count(reads.read_ptr[lo:hi + 1] - base, reads.read_tid[base:int(reads.read_ptr[hi])], reads.read_fraglen[lo:hi])
    s.counts_allreduce()
    dist.barrier()
    t0 = time.perf_counter()
    s.prepare(sharded=True)
    it, fd, ms = s.em_run(stop_on_conv=True)
    r = s.finalize()
    ctx.synchronize()
    sec = time.perf_counter() - t0
    st = s.model_stats()
    s.close()
    f = torch.from_numpy(r["fpkm"].copy()).cuda()
    f0 = f.clone()
    dist.broadcast(f0, 0)
    same = torch.tensor([1.0 if torch.equal(f, f0) else 0.0], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    tt = torch.tensor([sec, ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    pb = torch.tensor([float(st["peer_bytes_per_iter"]), float(st["bytes_per_iter"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(pb)
    out = None
    if rank == 0:
        s1 = ix.sample()
        s1.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
        r1 = s1.solve()
        st1 = s1.model_stats()
        s1.close()
        rel = np.abs(r["fpkm"] - r1["fpkm"]) / np.maximum(np.abs(r1["fpkm"]), 1e-300)
        rel = np.where(np.abs(r["ireadcount"] - r1["ireadcount"]) <= 1e-9, 0.0, rel)       # below 1e-9 reads the relative figure means nothing
        sec_all, ms_all = tt.tolist()
        peak, _ = peaks()
        one_us = 1e3 * r1["em_ms"] / max(r1["n_iter"], 1)
        us = 1e3 * ms_all / max(it, 1)
        out = {"workload": name, "n_gpus": world, "value": it / (ms_all * 1e-3), "unit": "iterations/s", "us_per_iter": us, "n_iter": int(it),
               "single_gpu_us_per_iter": one_us, "speedup_vs_one_gpu": one_us / us, "single_gpu_n_iter": int(r1["n_iter"]), "n_iter_diff": int(it - r1["n_iter"]),
               "parity_max_rel": float(rel.max()), "bitwise_equal_across_ranks": bool(same.item() == 1.0), "final_delta": float(fd),
               "nvlink_bytes_per_iter": int(pb[0].item()), "seconds_prepare_to_results": sec_all,
               "roofline_frac_of_n_gpus": (st1["bytes_per_iter"] / (us * 1e-6) / 1e9) / (peak * world),
               "exchange": "inside the EM kernel over NVLink peer memory (tagged 16-byte slots)" if ctx.comm_info()["peer_memory"] == 1 else "ncclAllReduce per iteration",
               "em_variant": st["em_variant"]}
    ix.close()
    return out


def note(msg):
    """Progress on stderr (rank and seconds since start): a leg that stalls is visible in the driver's log."""
    print(f"[bench r{os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


_T0 = time.perf_counter()
SELFTEST = "small"             # T = 20K, 200K classes, 3M reads: the small case that goes first


def sharded_child(args, rank, world, local):
    """`--leg class_sharded` (one process per GPU, started by the ranks of the main run): its own process group and contexts, so that a stall
    of the cross-GPU EM kernel at this GPU count can be cut off by the parent without losing the headline line."""
    fake = os.environ.get("EMSAR_BENCH_FAKE_LEG")          # tests/test_bench_cpu.py: the process plumbing without a GPU
    if fake:
        if fake == "hang":
            time.sleep(3600)
        if rank == 0:
            json.dump({"workload": args.sharded_workloads, "n_gpus": world, "fake": True}, open(args.leg_out, "w"))
        return
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from emsar_b200.api import Context
    ctx = Context(local)
    ctx.comm_init_torch()
    name = args.sharded_workloads
    note(f"sharded leg {name}: workload")
    sidx, sreads = bcast_workload(name, 1000, rank, torch, dist)
    note(f"sharded leg {name}: solve")
    o = class_sharded_leg(ctx, name, sidx, sreads, rank, world, torch, dist)
    note(f"sharded leg {name}: done")
    if rank == 0:
        with open(args.leg_out, "w") as f:
            json.dump(o, f)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def run_sharded_children(names, rank, world, local, budget_s, agree=None):
    """Every rank starts one child per workload (same RANK / WORLD_SIZE, the master port shifted) and waits for it at most budget_s[name]
    seconds; a child that is still running then is killed (by its pid). Rank 0 collects the children's results. `agree(failed)` returns
    whether the child failed on ANY rank (an all-reduce over the parents), so that all ranks take the same decision about going on."""
    import subprocess
    import tempfile
    out = []
    port = int(os.environ.get("MASTER_PORT", "29500"))
    for k, name in enumerate(names):
        leg_out = os.path.join(tempfile.gettempdir(), f"emsar_bench_leg_{port}_{k}.json")
        if rank == 0 and os.path.exists(leg_out):
            os.remove(leg_out)
        env = dict(os.environ, MASTER_PORT=str(port + 17 + k), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(local))
        if rank == 0:
            env["OMP_NUM_THREADS"] = str(min(16, os.cpu_count() or 1))       # rank 0 generates the workload; torchrun had set 1
        cmd = [sys.executable, os.path.abspath(__file__), "--leg", "class_sharded", "--gpus", str(world), "--sharded-workloads", name, "--leg-out", leg_out]
        note(f"class_sharded {name}: child started (limit {budget_s[name]} s)")
        pr = subprocess.Popen(cmd, env=env, stdout=subprocess.DEVNULL)
        _WATCH["children"].append(pr)
        try:
            rc = pr.wait(timeout=budget_s[name])
            err = None if rc == 0 else f"child exited with {rc}"
        except subprocess.TimeoutExpired:
            pr.kill()
            pr.wait()
            err = f"no result within {budget_s[name]} s at {world} GPUs (child stopped)"
        note(f"class_sharded {name}: {'ok' if err is None else err}")
        if agree is not None and agree(err is not None) and err is None:
            err = "the child of another rank did not come back"
        o = None
        if rank == 0:
            if err is None and os.path.exists(leg_out):
                o = json.load(open(leg_out))
                os.remove(leg_out)
            else:
                o = {"workload": name, "n_gpus": world, "error": err or "no result file"}
        out.append(o)
        if o is not None and "error" in o and name == SELFTEST:
            break                       # the small case did not come back: the full-size ones are not attempted
        if rank != 0 and err is not None and name == SELFTEST:
            break
    return out


_WATCH = {"line": None, "done": False, "children": []}


def start_watchdog(rank, limit_s):
    """A run that has not printed its line after limit_s seconds prints what it has (the headline without the legs that did not come back)
    and ends the process: a stalled leg must not turn into a run without a result."""
    def run():
        time.sleep(limit_s)
        if _WATCH["done"]:
            return
        note(f"watchdog: no result line after {limit_s} s")
        for pr in _WATCH["children"]:
            if pr.poll() is None:
                pr.kill()
        if rank == 0:
            line = _WATCH["line"]
            if line is not None:
                line = dict(line, incomplete=f"stopped by the bench watchdog after {limit_s} s: the legs that are missing did not come back")
                print(json.dumps(line), flush=True)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0 if _WATCH["line"] is not None or rank != 0 else 1)
    threading.Thread(target=run, daemon=True).start()


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
    ap.add_argument("--no-extras", action="store_true", help="tuning runs only: skip kernels / other_workloads / m64 / class_sharded / file_to_file")
    ap.add_argument("--others", default=",".join(OTHER_WORKLOADS), help="comma-separated workloads of the other_workloads object (N = 1)")
    ap.add_argument("--m64-per-gpu", type=int, default=8, help="samples per GPU of the -M batch leg (8 x 8 GPUs = BASELINE configs[3])")
    ap.add_argument("--sharded-workloads", default="config2_human_se,config5_full", help="N > 1: workloads of the class_sharded legs")
    ap.add_argument("--leg", default="", help="internal: run one leg as a child of the ranks of the main run (class_sharded)")
    ap.add_argument("--leg-out", default="", help="internal: where rank 0 of a child leg writes its result")
    ap.add_argument("--no-ref-binary", action="store_true",
                    help="skip timing the unmodified reference binary (oracle/_ref/emsar -p N) on the scaled twin (about 20 s)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.leg == "class_sharded":
        sharded_child(args, rank, world, local)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from emsar_b200.api import Context, Index

    note(f"start: {args.workload}, {world} GPU(s)")
    start_watchdog(rank, int(os.environ.get("EMSAR_BENCH_LIMIT_S", "1200")))
    idx, reads, gen_s = make_workload(args.workload, seed=1000 + rank)
    note("workload generated")
    ctx = Context(local)
    ix = Index(ctx, idx)
    # pinned host copies of this rank's read lists (the e2e leg copies them every step)
    # the compact wire form of the C ABI (emsar_sample_count_compact): uint16 lengths, int32 tids, uint16 fragment lengths (none for one length)
    one_fl = idx.nF == 1
    h_len = torch.from_numpy(np.diff(reads.read_ptr).astype(np.uint16)).pin_memory()
    h_tid = torch.from_numpy(reads.read_tid).pin_memory()
    h_fl = None if one_fl else torch.from_numpy(reads.read_fraglen.astype(np.uint16)).pin_memory()
    fl0 = int(reads.read_fraglen[0])
    h2d_bytes = h_len.numel() * 2 + h_tid.numel() * 4 + (0 if one_fl else h_fl.numel() * 2)

    def count(s_):
        s_.count_compact(h_len, h_tid, h_fl, fl0)

    def load(s_):
        count(s_)
        s_.prepare()

    smp = ix.sample()            # resident sample for the device-timed leg
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

    note("sample resident, warm-up")
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
    t_em = em_ms / 1e3           # device time of the EM kernel alone (CUDA events on the library's stream) -> roofline

    note("device-timed steps done")

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

    note("e2e steps done")
    # ---- one sample to convergence (samples/min context) ----
    conv = None
    if not args.no_converge:
        barrier()
        c0 = time.perf_counter()
        s = ix.sample()
        count(s)
        r = s.solve()
        s.close()
        c1 = time.perf_counter()
        conv_s = c1 - c0
        if world > 1:                                       # every rank solves a sample: the slowest decides
            tc = torch.tensor([conv_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            conv_s = float(tc.item())
        conv = {"seconds": conv_s, "samples_per_min": world * 60.0 / conv_s,
                "what": "host read lists -> counts -> model -> EM to convergence -> FPKM/TPM on the host, per sample", "n_iter": int(r["n_iter"]),
                "final_delta": float(r["final_delta"]), "em_ms": float(r["em_ms"]), "prep_ms": float(r["prep_ms"])}

    # ---- reduce over ranks: max time, summed work ----
    if world > 1:
        tt = torch.tensor([wall, e_wall, t_em], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall, e_wall, t_em = tt.tolist()
        ww = torch.tensor([iters_done, e_iters, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        iters_tot, e_iters_tot, launches_tot = ww.tolist()
    else:
        iters_tot, e_iters_tot, launches_tot = iters_done, e_iters, launches

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
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
    line = None
    n_reads_sample = len(reads.read_fraglen)
    if rank == 0:
        peak, peak_src = peaks()
        achieved = st["bytes_per_iter"] * iters_done / t_em / 1e9 if t_em > 0 else 0.0
        line = {
            "metric": "em_iterations_per_sec", "value": iters_tot / wall, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, idx, n_reads_sample, st["C_a"], st["nnz_a"], args.em_iters, world),
            "e2e": None if args.no_e2e else {"value": e_iters_tot / e_wall, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                                             "ms_per_step": 1e3 * e_wall / args.steps},
            "gpu_launches": int(launches_tot),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if world > 1 else measured_traffic(args.workload, args.em_iters),
                         "traffic_note": "DRAM bytes per launch (ncu, profiles/traffic.json); algorithmic bytes per launch = bytes_per_iter x "
                                         "em_iters_per_step: the packed model is L2-resident after the first iteration",
                         "algorithmic_bytes_per_launch": int(st["bytes_per_iter"]) * int(args.em_iters),
                         "peak_source": peak_src, "kernel": "k_em_psum" if st['em_variant'] == 5 else f"k_em_persistent<{st['em_variant']}>", "bytes_per_iter": st["bytes_per_iter"],
                         "us_per_iter": 1e6 * t_em / max(iters_done, 1)},
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "to_convergence": conv,
            "model": st, "gen_seconds": gen_s,
        }
        _WATCH["line"] = line                    # from here on the watchdog can still print the headline if a later leg stalls

    # ---- the other legs (outside every timed region above) ----
    extras = {}
    note("to-convergence sample done")
    if not args.no_extras:
        try:
            extras["m64"] = m64_leg(ctx, ix, idx, args.workload, rank, world, torch, dist, args.m64_per_gpu)
        except Exception as e:
            note(f"m64 leg: {e!r}")
            extras["m64"] = {"error": repr(e)}
        note("m64 leg done")
        if rank == 0:
            try:
                extras["kernels"] = [count_roofline(ctx, ix.sample, reads, torch)]
            except Exception as e:
                extras["kernels"] = [{"kernel": "k_count", "error": repr(e)}]
    smp.close()
    ix.close()
    del h_len, h_tid, h_fl
    if not args.no_extras and world > 1:
        # ONE sample over the N GPUs: in child processes (their own process group), a small case first; see run_sharded_children
        names = [SELFTEST] + [w for w in args.sharded_workloads.split(",") if w]
        budget = {SELFTEST: 150, "config2_human_se": 240, "config5_full": 360}
        if os.environ.get("EMSAR_BENCH_CHILD_LIMIT_S"):                       # tests: one short limit for every child
            budget = {n: int(os.environ["EMSAR_BENCH_CHILD_LIMIT_S"]) for n in names}
        torch.cuda.empty_cache()
        def agree(failed):
            fl = torch.tensor([1.0 if failed else 0.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(fl, op=dist.ReduceOp.MAX)
            return bool(fl.item() > 0)

        res = run_sharded_children(names, rank, world, local, {n: budget.get(n, 300) for n in names}, agree)
        dist.barrier()
        if rank == 0:
            extras["class_sharded"] = [o for o in res if o is not None]
    if not args.no_extras and world == 1:
        extras["other_workloads"] = []
        for name in [w for w in args.others.split(",") if w]:
            try:
                extras["other_workloads"].append(other_workload(ctx, name, torch, flush, args.em_iters))
            except Exception as e:
                extras["other_workloads"].append({"workload": name, "error": repr(e)})
        try:
            extras["file_to_file"] = file_to_file_twin(os.cpu_count() or 1, local) if not args.no_ref_binary else None
        except Exception as e:
            extras["file_to_file"] = {"error": repr(e)}
        note("file-to-file twin done")
        try:
            extras["index_build"] = index_build_leg(os.cpu_count() or 1, local)
        except Exception as e:
            extras["index_build"] = {"error": repr(e)}
        note("index construction leg done")

    if rank == 0:
        line.update(extras)
        _WATCH["done"] = True
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
