#!/usr/bin/env python
"""Where the host-buffer (e2e) step of bench.py spends its time: wall clock per C-ABI call, synchronised after each.
    python profiles/e2e_breakdown.py [workload] [em_iters]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from emsar_b200.api import Context, Index  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2_human_se"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    idx, reads, _ = bench.make_workload(wl, seed=1000)
    ctx = Context(0)
    ix = Index(ctx, idx)
    h_ptr = torch.from_numpy(reads.read_ptr).pin_memory()
    h_tid = torch.from_numpy(reads.read_tid).pin_memory()
    h_fl = torch.from_numpy(reads.read_fraglen).pin_memory()
    rows = []
    for rep in range(4):
        t = [time.perf_counter()]

        def mark():
            ctx.synchronize()
            t.append(time.perf_counter())
        s = ix.sample(); mark()
        s.count(h_ptr, h_tid, h_fl); mark()
        s.prepare(); mark()
        s.em_run(max_iter=iters, stop_on_conv=False); mark()
        s.finalize(); mark()
        s.close(); mark()
        rows.append([1e3 * (b - a) for a, b in zip(t[:-1], t[1:])])
    names = ["begin", "count(H2D+kernel)", "prepare", f"em_run({iters})", "finalize", "end"]
    for r in rows:
        print("  ".join(f"{n}={v:.2f}ms" for n, v in zip(names, r)), f" total={sum(r):.2f}ms")
    ix.close(); ctx.close()


if __name__ == "__main__":
    main()
