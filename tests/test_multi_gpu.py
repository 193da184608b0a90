"""Multi-GPU paths. On the GPU box with >= 2 devices: one sample class-range sharded over 2 ranks (NCCL all-reduce of the
per-transcript sums every iteration) must reproduce the single-GPU solve. On the CPU (gloo, world size 2): the host-side
logic of the same scheme — shard ranges, partial sums, all-reduce, replicated update — against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "fused_legacy", "fused_overflow", "nccl"])
def test_class_sharded_sample_two_gpus(built, mode):
    """fused: k_em_psum, the rows (and the classes of their median members) cut over the GPUs, shared rows exchanged over peer memory inside
    the kernel; fused_legacy / fused_overflow: the class-range sharded k_em_persistent<2> with its in-kernel all-reduce; nccl: ncclAllReduce
    per iteration."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(4, torch.cuda.device_count()) if mode in ("fused", "fused_legacy") else 2
    env = dict(os.environ)
    if mode == "nccl":
        env["EMSAR_SHARD_MODE"] = "nccl"
    if mode == "fused_overflow":
        env["EMSAR_EM_SMEM_KB"] = "12"
    if mode == "fused_legacy":
        env["EMSAR_EM_MODE"] = "legacy"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tests", "mgpu_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_shard_ranges_balanced(built):
    from emsar_b200.api import shard_ranges
    rng = np.random.default_rng(0)
    w = rng.integers(2, 100, size=5000)
    wp = np.concatenate([[0], np.cumsum(w)])
    for R in (1, 2, 3, 8):
        out = shard_ranges(wp, R)
        assert out[0] == 0 and out[-1] == len(w) and np.all(np.diff(out) >= 0)
        parts = [wp[out[r + 1]] - wp[out[r]] for r in range(R)]
        assert max(parts) - min(parts) <= 2 * w.max()


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT)
    from emsar_b200 import synth
    from emsar_b200.api import shard_ranges
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = synth.make_index(T=400, n_multi=2500, kmax=20, seed=41, module_cap=60)
    reads = synth.make_reads(idx, 20000, seed=41)
    R, F, N = oracle.count(idx, reads)
    Wf, adj, ps, iE = oracle.prepare(idx, F, N)
    T, cp, ct = idx.T, idx.class_ptr, idx.class_tid
    k = np.diff(cp)
    act = np.nonzero((np.arange(idx.C) >= T) & (R > 0) & (ps > 0))[0]          # active multi-tid classes, cid order
    wp = np.concatenate([[0], np.cumsum(k[act])])
    rng_ = shard_ranges(wp, world)
    mine = act[rng_[rank]:rng_[rank + 1]]                                       # this rank's class range
    A = np.zeros(T); np.add.at(A, ct, np.repeat(np.where(ps > 0, ps, 0.0), k))
    Rs = np.where(ps[:T] > 0, R[:T], 0).astype(float)
    theta = np.where(A > 0, 1.0, 0.0)
    n_steps = 8
    _, _, _, steps = oracle.em(idx, R, ps, None, max_iter=n_steps, n_steps=n_steps)
    err = 0.0
    for it in range(n_steps):
        Q = np.zeros(T)
        for c in mine:                                                          # E-step + partial M-step of the local classes
            m = ct[cp[c]:cp[c + 1]]
            s = theta[m].sum()
            if s > 0:
                np.add.at(Q, m, R[c] / s)
        tq = torch.from_numpy(Q)
        dist.all_reduce(tq)                                                     # the per-iteration all-reduce of the fp64 sums
        Q = tq.numpy()
        theta = np.where(A > 0, (Rs + theta * Q) / np.where(A > 0, A, 1.0), 0.0)
        err = max(err, float(np.max(np.abs(theta - steps[it]) / np.maximum(np.abs(steps[it]), 1e-300))))
    q.put((rank, err, int(len(mine))))
    dist.destroy_process_group()


def test_class_sharded_logic_gloo_world2(built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29547, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(e <= 1e-12 for _, e, _ in res), res
    assert all(n > 0 for _, _, n in res)


@pytest.mark.gpu
def test_cli_class_sharded_two_gpus(built, tmp_path):
    """`EMSAR_SHARD=classes EMSAR_DEVICES=0,1 emsar ...`: one sample on two GPUs (host threads of one process, peer access
    instead of CUDA IPC) must write the files the single-GPU run writes."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as gu
    fx = gu.FIXTURES["pe"]
    rsh, aln = gu.materialize(fx["rsh"], tmp_path), gu.materialize(fx["aln"], tmp_path)
    emsar = os.path.join(ROOT, "emsar_b200", "bin", "emsar")
    outs = {}
    for tag, env_extra in (("one", {}), ("two", {"EMSAR_DEVICES": "0,1", "EMSAR_SHARD": "classes"})):
        out = os.path.join(str(tmp_path), tag)
        env = dict(os.environ, **env_extra)
        r = subprocess.run([emsar, "-g", "-P", "-S", "-k", str(fx["k"]), "-s", fx["strand"], "-I", rsh, out, "p", aln], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[tag] = (out, r.stdout)
    assert "sharded by class range over 2 GPUs" in outs["two"][1]
    a, b = outs["one"][0], outs["two"][0]
    assert gu.parse_out_file(a + "/p.0.fraglength_effect") == gu.parse_out_file(b + "/p.0.fraglength_effect")
    sa, sb = gu.parse_out_file(a + "/p.0.segments"), gu.parse_out_file(b + "/p.0.segments")
    assert [r[:6] for r in sa] == [r[:6] for r in sb]
    fa, fb = gu.parse_out_file(a + "/p.0.fpkm"), gu.parse_out_file(b + "/p.0.fpkm")
    assert [r[0] for r in fa] == [r[0] for r in fb]
    for col in (1, 4, 6):           # FPKM, iReadcount, TPM: same optimum, sums in a different order
        x, y = np.array([float(r[col]) for r in fa]), np.array([float(r[col]) for r in fb])
        assert np.allclose(x, y, rtol=1e-6, atol=2e-6), col


@pytest.mark.gpu
def test_cli_multisample_over_two_gpus(built, tmp_path):
    """`EMSAR_DEVICES=0,1 emsar -M ...`: the files of the list are spread over the GPUs (one host thread each, no communication)
    and every output file is the one the single-GPU run writes, byte for byte."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as gu
    fx = gu.FIXTURES["se"]
    rsh, aln = gu.materialize(fx["rsh"], tmp_path), gu.materialize(fx["aln"], tmp_path)
    lines = open(aln).read().splitlines(True)
    files = [aln]
    for i, frac in enumerate((2, 3, 4)):
        p = os.path.join(str(tmp_path), f"part{i}.bowtie")
        open(p, "w").writelines(lines[:len(lines) // frac])
        files.append(p)
    lst = os.path.join(str(tmp_path), "list.txt")
    open(lst, "w").write("\n".join(files) + "\n")
    emsar = os.path.join(ROOT, "emsar_b200", "bin", "emsar")
    outs = []
    for tag, env_extra in (("one", {}), ("two", {"EMSAR_DEVICES": "0,1"})):
        out = os.path.join(str(tmp_path), tag)
        r = subprocess.run([emsar, "-q", "-g", "-M", "-I", rsh, out, "p", lst], capture_output=True, text=True, env=dict(os.environ, **env_extra), timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(out)
    for i in range(len(files)):
        for ext in ("fpkm", "fraglength_effect", "segments"):
            a, b = (open(os.path.join(o, f"p.{i}.{ext}")).read() for o in outs)
            assert a == b, (i, ext)
