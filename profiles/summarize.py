#!/usr/bin/env python
"""Turns gpurun_out/launches_*.csv (ncu --metrics gpu__time_duration.sum) and gpurun_out/prof_*.ncu-rep (ncu --set full)
into the small text summaries committed under profiles/.   usage: summarize.py <tag> <launches.csv> <prof.ncu-rep>"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "inst_executed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__shared_mem_per_block_static", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        a = agg.setdefault(r[ki][:70], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    out.write("kernel                                                                  launches   total_us   share\n")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.write(f"{n:70s} {c:8d} {t / 1e3:10.1f} {100 * t / tot:6.1f}%\n")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    for v in rows[2:]:
        out.write(f"\n== {v[H.index('Kernel Name')]}  (launch id {v[0]})\n")
        for k in KEYS:
            if k in H:
                out.write(f"{k:80s} {U[H.index(k)]:>12s} {v[H.index(k)]}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    H, data = rows[1], rows[2:]
    isrc, ismp, iex = H.index("Source"), H.index("# Samples"), H.index("Instructions Executed")
    stall = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    ts, te = sum(int(r[ismp]) for r in data), sum(int(r[iex]) for r in data)
    out.write(f"\nSASS hot spots (of {ts} samples, {te} warp instructions)\n")
    for n, r in enumerate(data):
        s, e = int(r[ismp]), int(r[iex])
        if s > ts * 0.01:
            st = sorted([(int(r[i]), H[i][6:]) for i in stall if r[i] not in ("", "0")], reverse=True)[:2]
            out.write(f"{n:5d} {r[isrc].strip()[:56]:56s} samples {100 * s / ts:5.1f}%  exec {100 * e / te:5.1f}%  {st}\n")
    tot = collections.Counter()
    for r in data:
        for i in stall:
            if r[i] not in ("", "0"):
                tot[H[i][6:]] += int(r[i])
    out.write("\nstall reasons over all samples: " + ", ".join(f"{k} {100 * v / max(ts, 1):.1f}%" for k, v in tot.most_common(8)) + "\n")


if __name__ == "__main__":
    tag, lcsv, rep = sys.argv[1:4]
    with open(f"profiles/{tag}_launches.txt", "w") as f:
        launches(lcsv, f)
    with open(f"profiles/{tag}_em_kernel.txt", "w") as f:
        full(rep, f)
    print("wrote profiles/%s_*.txt" % tag)
