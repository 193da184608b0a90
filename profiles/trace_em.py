import sys, ctypes as C, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from emsar_b200.api import Context, Index
from emsar_b200 import _lib
idx, reads, _ = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "config2_human_se", 1000)
ctx = Context(0); ix = Index(ctx, idx); s = ix.sample()
s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen); s.prepare()
s.em_run(max_iter=50, stop_on_conv=False)
out = np.zeros(148*8 + 64 + 1600, dtype=np.uint64); nb = C.c_int(0)
rc = _lib.lib().emsar_debug_em_trace(s._h, 30, out.ctypes.data_as(C.c_void_p), C.byref(nb))
t = out.reshape(-1,8)[:nb.value,:5].astype(np.int64)
t0 = t[:,0].min()
E = t[:,1]-t[:,0]; W1 = t[:,2]-t[:,1]; M = t[:,3]-t[:,2]; W2 = t[:,4]-t[:,3]
print("rc", rc, "blocks", nb.value)
for n,a in [("E dur",E),("wait1",W1),("M dur",M),("wait2",W2)]:
    print(f"{n:6s} min {a.min()/1e3:7.2f} mean {a.mean()/1e3:7.2f} max {a.max()/1e3:7.2f} us")
print("E start spread", (t[:,0].max()-t0)/1e3, "iter total", (t[:,4].max()-t0)/1e3)
print("barrier1 release spread", (t[:,2].max()-t[:,2].min())/1e3, "last E end -> first release", (t[:,2].min()-t[:,1].max())/1e3)
print("barrier2: last M end -> first release", (t[:,4].min()-t[:,3].max())/1e3)


order = np.argsort(-M)
print("slowest M CTAs:", [(int(i), round(M[i]/1e3,2)) for i in order[:5]], "fastest:", [(int(i), round(M[i]/1e3,2)) for i in order[-3:]])
order = np.argsort(-E)
print("slowest E CTAs:", [(int(i), round(E[i]/1e3,2)) for i in order[:5]], "fastest:", [(int(i), round(E[i]/1e3,2)) for i in order[-3:]])

tr = out[nb.value*8+64:nb.value*8+64+1600].astype(np.int64).reshape(-1,4)
tr = tr[tr[:,0]>0]
if len(tr):
    b0 = t[0,0]
    k = tr[:,2] & 0xffff; mode=(tr[:,2]>>16)&0xff; res=(tr[:,2]>>40)&1; w=(tr[:,2]>>48)
    dur = (tr[:,1]-tr[:,0])/1e3; st=(tr[:,0]-b0)/1e3
    print("CTA 0 E tiles:", len(tr), "sum dur", dur.sum().round(1), "us; per warp busy mean", (dur.sum()/32).round(2))
    for kk in sorted(set(k.tolist())):
        m = k==kk
        print(f"  k={kk:3d} mode {mode[m][0]} tiles {m.sum():3d} resident {res[m].sum():3d} dur mean {dur[m].mean():5.2f} max {dur[m].max():5.2f} us  start {st[m].min():5.2f}..{st[m].max():5.2f} cnt {tr[m,3].mean():5.1f}")
    print("  last tile end", ((tr[:,1]-b0)/1e3).max().round(2))
