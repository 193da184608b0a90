"""Worker of the multi-GPU tests (launched with torch.distributed.run, one process per GPU):
one sample, class-range sharded over the ranks, against the single-GPU solve of the same sample on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emsar_b200 import synth  # noqa: E402
from emsar_b200.api import Context, Index  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    mode = sys.argv[1] if len(sys.argv) > 1 else "fused"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idx = synth.make_index(T=3000, n_multi=20000, alpha=1.8, kmax=120, seed=31, module_cap=400)
    reads = synth.make_reads(idx, 300000, seed=31)
    ctx = Context(local)
    ctx.comm_init_torch()
    ix = Index(ctx, idx)
    # every rank counts its own slice of the read groups; the integer counts are summed over the ranks
    n = len(reads.read_fraglen)
    lo, hi = n * rank // world, n * (rank + 1) // world
    s = ix.sample()
    s.count(reads.read_ptr[lo:hi + 1], reads.read_tid, reads.read_fraglen[lo:hi])
    s.counts_allreduce()
    R, F, N = s.counts()
    r = s.solve(sharded=True)
    st = s.model_stats()
    s.close()
    ok = True
    msg = ""
    pm = ctx.comm_info()["peer_memory"]
    if pm != (-1 if mode == "nccl" else 1):
        ok, msg = False, f"expected the {mode} path, peer_memory={pm}"
    # reference: the whole sample on one GPU (every rank does it: also checks that all ranks hold the same answer)
    s1 = ix.sample()
    s1.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
    R1, F1, N1 = s1.counts()
    r1 = s1.solve()
    st1 = s1.model_stats()
    s1.close()
    if not (N == N1 and np.array_equal(R, R1) and np.array_equal(F, F1)):
        ok, msg = False, "sharded counts differ"
    rel = np.abs(r["fpkm"] - r1["fpkm"]) / np.maximum(np.abs(r1["fpkm"]), 1e-300)
    absr = np.abs(r["ireadcount"] - r1["ireadcount"])
    if not np.all((rel <= 1e-9) | (absr <= 1e-9)):
        ok, msg = False, f"sharded fpkm differs: {rel.max()}"
    if abs(r["n_iter"] - r1["n_iter"]) > 1:
        ok, msg = False, f"iterations {r['n_iter']} vs {r1['n_iter']}"
    # bit-identical across ranks
    t = torch.from_numpy(r["fpkm"].copy()).cuda()
    t0 = t.clone()
    dist.broadcast(t0, 0)
    if not torch.equal(t, t0):
        ok, msg = False, "ranks disagree bitwise"
    tot = torch.tensor([float(st["nnz_a"])], device="cuda")
    dist.all_reduce(tot)
    if st["em_variant"] == 5:
        # k_em_psum: every rank packs the whole model and runs its own CTAs of it (the rows are cut over the ranks, not the classes)
        if st["nnz_a"] != st1["nnz_a"] or (world > 1 and st["peer_bytes_per_iter"] <= 0):
            ok, msg = False, f"psum sharding: nnz_a {st['nnz_a']} vs {st1['nnz_a']}, peer bytes {st['peer_bytes_per_iter']}"
    elif int(tot.item()) != st1["nnz_a"]:
        ok, msg = False, f"shards do not partition the active classes: {int(tot.item())} vs {st1['nnz_a']}"
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"MGPU {'OK' if flag.item() == 1.0 else 'FAIL'} mode={mode} variant={st['em_variant']} peer_memory={pm} world={world} iters={r['n_iter']}/{r1['n_iter']} em_ms={r['em_ms']:.1f}/{r1['em_ms']:.1f} "
              f"max_rel={float(rel.max()):.2e} nnz_a/rank={st['nnz_a']} {msg}", flush=True)
    ix.close(); ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
