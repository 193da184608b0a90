#!/usr/bin/env python
"""Per-phase timing of the fused class-sharded EM kernel (globaltimer stamps of the last iteration, per CTA).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/trace_shard.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from emsar_b200 import _lib  # noqa: E402
from emsar_b200.api import Context, Index  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idx, reads, _ = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "config2_human_se", 1000)
ctx = Context(local)
ctx.comm_init_torch()
ix = Index(ctx, idx)
s = ix.sample()
s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
s.prepare(sharded=True)
s.em_run(max_iter=50, stop_on_conv=False)
out = np.zeros(148 * 8 + 64 + 1600, dtype=np.uint64)
nb = C.c_int(0)
rc = _lib.lib().emsar_debug_em_trace(s._h, 30, out.ctypes.data_as(C.c_void_p), C.byref(nb))
t = out.reshape(-1, 8)[:nb.value, :7].astype(np.int64)
names = ["theta wait + E", "delta(it-1) read + grid barrier", "M partial + push", "owner update (waits for partials) + push", "delta publish"]
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank} rc {rc} peer_memory {ctx.comm_info()['peer_memory']} iteration total {(t[:, 5].max() - t[:, 0].min()) / 1e3:.2f} us")
        for i, n in enumerate(names):
            a = (t[:, i + 1] - t[:, i]) / 1e3
            print(f"  {n:42s} min {a.min():6.2f} mean {a.mean():6.2f} max {a.max():6.2f} us")
        sys.stdout.flush()
s.close(); ix.close(); ctx.close()
dist.destroy_process_group()
