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
