#!/usr/bin/env python
"""The HBM-bound one-off of config #3 (PE index, nF = 400 fragment lengths): adjEUMA[c] = sum_i Wf[i] * EUMA[c][i] streams
4*C*nF bytes (3.2 GB at 2M classes). Prints the model-build time; run under
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:k_adjeuma
for the kernel's own duration and traffic.   usage: adjeuma_bench.py [n_multi] [nF]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emsar_b200 import synth  # noqa: E402
from emsar_b200.api import Context, Index  # noqa: E402

n_multi = int(sys.argv[1]) if len(sys.argv) > 1 else 1_900_000
nF = int(sys.argv[2]) if len(sys.argv) > 2 else 400
t0 = time.time()
idx = synth.make_index(T=200000, n_multi=n_multi, alpha=2.4, kmax=99, seed=3, module_cap=5000, nF=nF, frag_min=101, readlength=101)
reads = synth.make_reads(idx, 2_000_000, seed=3)
print(f"index T={idx.T} C={idx.C} nF={idx.nF}: EUMA {4 * idx.C * idx.nF / 1e9:.2f} GB, generated in {time.time() - t0:.0f} s", flush=True)
ctx = Context(0)
ix = Index(ctx, idx)
s = ix.sample()
s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
for rep in range(4):
    t0 = time.perf_counter()
    s.prepare()
    ctx.synchronize()
    print(f"prepare {1e3 * (time.perf_counter() - t0):.2f} ms")
s.close(); ix.close(); ctx.close()
