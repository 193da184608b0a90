"""Per-CTA globaltimer stamps of one iteration of k_em_psum (emsar_debug_em_trace): halo wait / E / convergence read + M / U.
usage: python profiles/trace_psum.py [workload]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from emsar_b200 import _lib
from emsar_b200.api import Context, Index

name = sys.argv[1] if len(sys.argv) > 1 else "config2_human_se"
idx, reads, _ = bench.make_workload(name, 1000)
ctx = Context(0); ix = Index(ctx, idx); s = ix.sample()
s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen); s.prepare()
st = s.model_stats()
s.em_run(max_iter=50, stop_on_conv=False)
out = np.zeros(148 * 8 + 64 + 1600, dtype=np.uint64); nb = C.c_int(0)
rc = _lib.lib().emsar_debug_em_trace(s._h, 30, out.ctypes.data_as(C.c_void_p), C.byref(nb))
t = out.reshape(-1, 8)[:nb.value, :5].astype(np.int64)
t0 = t[:, 0].min()
H = t[:, 1] - t[:, 0]; E = t[:, 2] - t[:, 1]; M = t[:, 3] - t[:, 2]; U = t[:, 4] - t[:, 3]
print(name, "variant", st["em_variant"], "rc", rc, "blocks", nb.value)
for n, a in [("halo wait", H), ("E", E), ("dm read + M", M), ("U (partials, update)", U), ("whole", t[:, 4] - t[:, 0])]:
    print(f"{n:22s} min {a.min() / 1e3:7.2f} mean {a.mean() / 1e3:7.2f} max {a.max() / 1e3:7.2f} us")
print("start spread", (t[:, 0].max() - t0) / 1e3, "iteration (first start -> last end)", (t[:, 4].max() - t0) / 1e3)
for lab, a in (("E", E), ("M", M), ("U", U), ("halo", H)):
    o = np.argsort(-a)
    print("slowest", lab, [(int(i), round(a[i] / 1e3, 2)) for i in o[:5]], "fastest", [(int(i), round(a[i] / 1e3, 2)) for i in o[-3:]])

# per-CTA table for cost-model work: features of the CTA's share of the packed model (emsar_debug_psum_blocks) next to its phase times
feat = np.zeros(nb.value * 16, dtype=np.int32); nb2 = C.c_int(0)
fn = getattr(_lib.lib(), "emsar_debug_psum_blocks", None)
if fn is not None and st["em_variant"] == 5 and fn(s._h, feat.ctypes.data_as(C.c_void_p), C.byref(nb2)) == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    path = f"gpurun_out/psum_blocks_{name}{os.environ.get('TRACE_TAG', '')}.csv"
    with open(path, "w") as f:
        f.write("cta,rows,halo,classes,incoming,etiles,mitems,res16,desc,k2_tiles,k34_tiles,lane_steps,multi_steps,res_etiles,slice_steps,group_blocks,res_mitems,halo_ns,E_ns,M_ns,U_ns\n")
        for i in range(nb2.value):
            f.write(",".join([str(i)] + [str(int(x)) for x in feat[16 * i:16 * i + 16]] + [str(int(H[i])), str(int(E[i])), str(int(M[i])), str(int(U[i]))]) + "\n")
    print("per-CTA table:", path)
