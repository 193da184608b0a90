"""bench.py's multi-GPU plumbing without a GPU: the class_sharded legs run in child processes that the ranks of the main run start and cut
off after a time limit (a stall of the cross-GPU EM kernel must not cost the headline line). The children are faked through
EMSAR_BENCH_FAKE_LEG (ok = write a result, hang = never return)."""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _rank(rank, world, mode, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_PORT=str(port), MASTER_ADDR="127.0.0.1", EMSAR_BENCH_FAKE_LEG=mode)
    import bench
    t0 = time.perf_counter()
    res = bench.run_sharded_children([bench.SELFTEST, "config2_human_se"], rank, world, rank, {bench.SELFTEST: 4, "config2_human_se": 4})
    q.put((rank, res, time.perf_counter() - t0))


def _run(mode, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rank, args=(r, 2, mode, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=30)
    return got


def test_children_report_through_rank0():
    (r0, res0, _), (r1, res1, _) = _run("ok", 29811)
    assert [o["workload"] for o in res0] == ["small", "config2_human_se"] and all(o["fake"] for o in res0)
    assert res1 == [None, None]


def test_a_stalled_child_is_cut_off_and_the_big_legs_are_skipped():
    (r0, res0, dt0), (r1, res1, dt1) = _run("hang", 29841)
    assert len(res0) == 1 and "no result within 4 s" in res0[0]["error"]        # the small case went first and did not come back
    assert res1 == [None]
    assert dt0 < 60 and dt1 < 60
    json.dumps(res0)


# ---- the single-GPU control flow of bench.main() against stand-ins for the device API (no CUDA in this container) -------------------
class _FakeSample:
    def __init__(self, idx):
        self.idx, self.n = idx, 0

    def count_compact(self, *a): pass
    def count(self, *a): pass
    def prepare(self, **k): pass
    def close(self): pass
    def time_adjeuma(self, reps=5): return 0.5

    def model_stats(self):
        return {"T": self.idx.T, "C_a": 10, "nnz_a": 30, "bytes_per_iter": 8 * 30 + 24 * 10 + 44 * self.idx.T, "em_variant": 5, "all_local": 1,
                "index_bytes": 100, "peer_bytes_per_iter": 0}

    def em_run(self, max_iter=0, stop_on_conv=True, reset_theta=False):
        return (max_iter or 123), 0.9, 1.0

    def _res(self):
        import numpy as np
        T = self.idx.T
        return {"fpkm": np.ones(T), "tpm": np.full(T, 1e6 / T), "ireadcount": np.ones(T), "n_iter": 123, "final_delta": 0.9, "em_ms": 1.0, "prep_ms": 0.1}

    def finalize(self): return self._res()
    def solve(self, **k): return self._res()

    def counts(self):
        import numpy as np
        C = len(self.idx.class_ptr) - 1
        return np.ones(C, dtype=np.int32), np.ones(self.idx.max_fraglength + 1, dtype=np.int32), C


class _FakeIndex:
    def __init__(self, ctx, idx): self.idx = idx
    def sample(self): return _FakeSample(self.idx)
    def close(self): pass


class _FakeContext:
    def __init__(self, dev=0): self.k = 0
    def launches(self): self.k += 7; return self.k
    def synchronize(self): pass
    def timer_start(self): pass
    def timer_stop(self): return 1.0
    def close(self): pass


def test_single_gpu_control_flow_prints_one_complete_line(monkeypatch, capsys):
    import torch
    import bench
    from emsar_b200 import api
    monkeypatch.setattr(api, "Context", _FakeContext)
    monkeypatch.setattr(api, "Index", _FakeIndex)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    real_empty, real_tensor = torch.empty, torch.tensor
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "device"}))
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != "device"}))
    monkeypatch.setattr(bench, "file_to_file_twin", lambda cores, dev: {"fake": True})
    real_leg = bench.index_build_leg
    monkeypatch.setattr(bench, "index_build_leg", lambda cores, dev: real_leg(cores, dev, transcripts=400))
    monkeypatch.setattr(bench, "OTHER_WORKLOADS", ["tiny"])
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "tiny", "--steps", "2", "--warmup", "1", "--em-iters", "5", "--m64-per-gpu", "1", "--others", "tiny"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    # the -M leg draws 20-40M reads per sample: shrink it
    real_make_reads = bench.make_reads
    monkeypatch.setattr(bench, "make_reads", lambda name, idx, seed, n_reads=None: real_make_reads(name, idx, seed, n_reads=None if n_reads is None else 1000))
    bench.main()
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "to_convergence", "m64", "kernels", "other_workloads", "file_to_file", "index_build"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert "error" not in line["m64"] and "error" not in line["kernels"][0] and "error" not in line["other_workloads"][0], line
    ib = line["index_build"]                 # no GPU here: the device arms fail loudly, the host builder and the reference agree with each other
    assert "no CUDA device" in ib["device"]["error"] and ib["host"]["classes"] > 100 and "sha256" in ib["reference_quarter"]
    assert "incomplete" not in line


def _mock_everything():
    """the same stand-ins, applied by hand inside a spawned rank (no monkeypatch fixture there)"""
    import torch
    import torch.distributed as dist
    import bench
    from emsar_b200 import api
    api.Context, api.Index = _FakeContext, _FakeIndex
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.pin_memory = lambda self: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    real_empty, real_tensor, real_init = torch.empty, torch.tensor, dist.init_process_group
    torch.empty = lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "device"})
    torch.tensor = lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != "device"})
    dist.init_process_group = lambda backend, device_id=None: real_init("gloo")
    real_make_reads = bench.make_reads
    bench.make_reads = lambda name, idx, seed, n_reads=None: real_make_reads(name, idx, seed, n_reads=None if n_reads is None else 1000)
    return bench


def _main_rank(rank, world, port, path, mode="ok"):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_PORT=str(port), MASTER_ADDR="127.0.0.1", EMSAR_BENCH_FAKE_LEG=mode,
                      EMSAR_BENCH_CHILD_LIMIT_S="5")
    bench = _mock_everything()
    sys.argv = ["bench.py", "--gpus", str(world), "--workload", "tiny", "--steps", "2", "--warmup", "1", "--em-iters", "5", "--m64-per-gpu", "1"]
    if rank == 0:
        sys.stdout = open(path, "w")
    bench.main()
    sys.stdout.flush()


def test_two_rank_control_flow_over_gloo(tmp_path):
    """world_size 2 on the CPU: barriers, reductions over the ranks, the -M leg, the child legs, one line from rank 0."""
    ctx = mp.get_context("spawn")
    path = str(tmp_path / "rank0.out")
    ps = [ctx.Process(target=_main_rank, args=(r, 2, 29871, path)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(timeout=300)
        assert p.exitcode == 0
    out = [l for l in open(path).read().splitlines() if l.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["cpu_baseline"] is None
    assert line["m64"]["samples"] == 2 and "error" not in line["m64"]
    assert [o["workload"] for o in line["class_sharded"]] == ["small", "config2_human_se", "config5_full"]
    assert "incomplete" not in line


def test_two_rank_run_survives_a_child_that_never_returns(tmp_path):
    ctx = mp.get_context("spawn")
    path = str(tmp_path / "rank0.out")
    ps = [ctx.Process(target=_main_rank, args=(r, 2, 29901, path, "hang")) for r in range(2)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    for p in ps:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert time.perf_counter() - t0 < 120
    line = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][0])
    assert line["n_gpus"] == 2 and line["value"] > 0 and "error" not in line["m64"]
    assert len(line["class_sharded"]) == 1 and "no result within 5 s" in line["class_sharded"][0]["error"]
