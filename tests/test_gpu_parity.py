"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Integer results are compared bit-exact; fp64 results within the tolerances written next to each assert."""
import numpy as np
import pytest

from emsar_b200 import synth
from emsar_b200.api import Index

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import oracle
    return oracle


CASES = {
    "se_small": dict(idx=dict(T=300, n_multi=1500, kmax=12, seed=1, module_cap=40), N=20000),
    "se_longk": dict(idx=dict(T=2000, n_multi=8000, alpha=1.5, kmax=300, seed=2, module_cap=400), N=60000),
    "pe_nf21": dict(idx=dict(T=500, n_multi=3000, kmax=20, seed=3, module_cap=60, nF=21, frag_min=40, readlength=25), N=30000),
    "no_node": dict(idx=dict(T=400, n_multi=2000, kmax=9, seed=4, module_cap=50, p_no_node=0.2), N=20000),
    "hubs": dict(idx=dict(T=3000, n_multi=6000, kmax=30, seed=5, module_cap=3000, hubs=3, hub_classes=6000), N=200000),
}


def _make(name):
    c = CASES[name]
    idx = synth.make_index(**c["idx"])
    reads = synth.make_reads(idx, c["N"], seed=c["idx"]["seed"])
    return idx, reads


@pytest.mark.parametrize("name", list(CASES))
def test_count_bit_exact(ctx, name):
    idx, reads = _make(name)
    # reads outside the fragment-length window are dropped entirely (:849)
    reads.read_fraglen[::97] = idx.max_fraglength + 5
    R0, F0, N0 = _oracle().count(idx, reads)
    ix = Index(ctx, idx)
    s = ix.sample()
    # two batches: counts accumulate across calls like successive update_ReadCounts calls
    n = len(reads.read_fraglen)
    h = n // 3
    s.count(reads.read_ptr[:h + 1], reads.read_tid, reads.read_fraglen[:h])
    s.count(reads.read_ptr[h:], reads.read_tid, reads.read_fraglen[h:])
    R, F, N = s.counts()
    s.close(); ix.close()
    assert N == N0
    assert np.array_equal(F, F0)
    assert np.array_equal(R, R0)
    assert R.sum() > 0


@pytest.mark.parametrize("name", list(CASES))
def test_solve_matches_oracle(ctx, name):
    idx, reads = _make(name)
    o = _oracle().quantify(idx, reads)
    ix = Index(ctx, idx)
    s = ix.sample()
    s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
    r = s.solve()
    adj, ex, cs = s.segments()
    wf = s.wf()
    st = s.model_stats()
    s.close(); ix.close()
    # deterministic fp64 pre-steps: bit-identical (same summation order, no FMA contraction)
    assert np.array_equal(wf, o["Wf"])
    assert np.array_equal(adj, o["adjEUMA"])
    assert np.array_equal(r["efflen"], o["iEUMA"])
    assert np.array_equal(cs, o["CS"])
    assert r["total_readcount"] == o["N"]
    # estimator: same update, same stopping rule -> iteration count within +-1, outputs <= 1e-9 relative
    assert abs(r["n_iter"] - o["n_iter"]) <= 1, (r["n_iter"], o["n_iter"])
    assert r["final_delta"] <= 1.0
    scale = np.maximum(np.abs(o["fpkm"]), 1e-300)
    rel = np.abs(r["fpkm"] - o["fpkm"]) / scale
    reads_abs = np.abs(r["ireadcount"] - o["ireadcount"])
    assert np.all((rel <= 1e-9) | (reads_abs <= 1e-9)), (rel.max(), reads_abs.max())
    assert np.allclose(r["tpm"], o["tpm"], rtol=1e-9, atol=1e-9)
    assert np.allclose(ex, o["expected"], rtol=1e-9, atol=1e-9)
    assert abs(r["loglik"] - o["loglik"]) <= 1e-10 * abs(o["loglik"])
    assert np.array_equal(r["ireadcount_int"], o["ireadcount_int"]) or np.abs(r["ireadcount"] - o["ireadcount"]).max() < 1e-6
    assert st["C_a"] == int(((o["ReadCount"][idx.T:] > 0) & (o["EUMAps"][idx.T:] > 0)).sum())


@pytest.mark.parametrize("name", ["se_small", "se_longk", "hubs"])
def test_em_iterates_match_oracle(ctx, name):
    """theta after each of the first iterations: <= 1e-12 relative."""
    idx, reads = _make(name)
    orc = _oracle()
    R, F, N = orc.count(idx, reads)
    Wf, adj, ps, iE = orc.prepare(idx, F, N)
    n_steps = 12
    _, _, _, steps = orc.em(idx, R, ps, None, max_iter=n_steps, n_steps=n_steps)
    ix = Index(ctx, idx)
    s = ix.sample()
    s.set_counts(R, F)
    s.prepare()
    for it in range(n_steps):
        s.em_run(max_iter=1, stop_on_conv=False)
        th = s.theta()
        ref = steps[it]
        err = np.abs(th - ref) / np.maximum(np.abs(ref), 1e-300)
        assert err.max() <= 1e-12, (it, err.max())
    s.close(); ix.close()


def test_eumacut_loop(ctx):
    """Sets larger than the cap raise EUMAcut by 2 until they fit (emsar_main.c:411-425)."""
    idx = synth.make_index(T=600, n_multi=4000, kmax=10, seed=11, module_cap=300)
    reads = synth.make_reads(idx, 30000, seed=11)
    orc = _oracle()
    R, F, N = orc.count(idx, reads)
    Wf, adj, ps, iE = orc.prepare(idx, F, N)
    max_sid, cut, CS, TS = orc.components(idx, adj, 0.0, max_ntid=50)
    assert cut > 0 and (CS < 0).any()
    in_model = (CS >= 0).astype(np.uint8)
    th, n_iter, fd, _ = orc.em(idx, R, ps, in_model)
    ix = Index(ctx, idx)
    s = ix.sample()
    s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
    r = s.solve(max_ntid_per_sid=50)
    _, _, cs = s.segments()
    s.close(); ix.close()
    assert r["eumacut"] == cut
    assert np.array_equal(cs, CS)
    assert r["max_sid"] == max_sid
    assert abs(r["n_iter"] - n_iter) <= 1
    assert np.allclose(r["fpkm"], th, rtol=1e-9, atol=1e-12)


def test_empty_and_zero_read_samples(ctx):
    idx = synth.make_index(T=100, n_multi=300, kmax=6, seed=12, module_cap=20)
    ix = Index(ctx, idx)
    s = ix.sample()
    s.count(np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))   # empty batch
    R, F, N = s.counts()
    assert N == 0 and R.sum() == 0
    r = s.solve()          # no reads at all: every FPKM is 0 (MLE :3054-3059); Wf is 0/0 like the reference
    assert np.all(r["fpkm"] == 0)
    s.close(); ix.close()


def test_unsupported_and_bad_inputs(ctx):
    from emsar_b200._lib import EmsarError
    idx = synth.make_index(T=100, n_multi=300, kmax=6, seed=13, module_cap=20)
    ix = Index(ctx, idx)
    s = ix.sample()
    k = 1025
    s.count(np.array([0, k], dtype=np.int64), np.zeros(k, dtype=np.int32), np.ones(1, dtype=np.int32))
    with pytest.raises(EmsarError):
        s.counts()
    s.close()
    s = ix.sample()
    with pytest.raises(EmsarError):
        s.solve()          # no counts yet
    s.close(); ix.close()
    bad = synth.make_index(T=100, n_multi=300, kmax=6, seed=13, module_cap=20)
    bad.class_tid = bad.class_tid.copy()
    bad.class_tid[bad.class_ptr[150]:bad.class_ptr[151]] = bad.class_tid[bad.class_ptr[150]:bad.class_ptr[151]][::-1] + 0
    if not np.array_equal(bad.class_tid, idx.class_tid):
        with pytest.raises(EmsarError):
            Index(ctx, bad)


# The EM kernel has four variants: the barrier-free one (everything of every CTA resident in shared memory: the usual case
# at test sizes and at config #2), the same with grid barriers, the overflow path (halo rows / classes and q that did not get a slot are read from the L2-resident global copies) and the
# TMA-pipelined index streams. The two rarer ones are forced here with the tuning knobs the library reads from the environment.
VARIANTS = {
    "overflow_12k": {"EMSAR_EM_SMEM_KB": "12"},
    "overflow_24k": {"EMSAR_EM_SMEM_KB": "24"},
    "grid_barriers": {"EMSAR_EM_MODE": "barrier"},          # the default at these sizes is the barrier-free kernel
    "pipelined": {"EMSAR_EM_MODE": "pipe"},
    "pipelined_small": {"EMSAR_EM_MODE": "pipe", "EMSAR_EM_SMEM_KB": "80"},
}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("name", ["se_longk", "hubs"])
def test_solve_kernel_variants(built, monkeypatch, variant, name):
    from emsar_b200.api import Context
    for k, v in VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    idx, reads = _make(name)
    o = _oracle().quantify(idx, reads)
    c = Context(0)
    try:
        ix = Index(c, idx)
        s = ix.sample()
        s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
        r = s.solve()
        s.close(); ix.close()
    finally:
        c.close()
    assert abs(r["n_iter"] - o["n_iter"]) <= 1, (r["n_iter"], o["n_iter"])
    rel = np.abs(r["fpkm"] - o["fpkm"]) / np.maximum(np.abs(o["fpkm"]), 1e-300)
    reads_abs = np.abs(r["ireadcount"] - o["ireadcount"])
    assert np.all((rel <= 1e-9) | (reads_abs <= 1e-9)), (rel.max(), reads_abs.max())


def test_short_launch_then_normal_sample_same_context(built):
    """A launch that ends after 1-2 iterations leaves low tags in the convergence slots of the barrier-free kernel; the next
    sample on the SAME context restarts its tags at 0 and must not accept them (the slots are cleared at every launch)."""
    from emsar_b200.api import Context
    idx, reads = _make("se_longk")
    o = _oracle().quantify(idx, reads)
    c = Context(0)
    try:
        ix = Index(c, idx)
        for short in (1, 2, 3, 2, 1):
            a = ix.sample()
            a.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
            a.prepare()
            it, fd, ms = a.em_run(max_iter=short, stop_on_conv=True)
            assert it == short
            a.close()
            s = ix.sample()
            s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
            r = s.solve()
            s.close()
            assert abs(r["n_iter"] - o["n_iter"]) <= 1, (short, r["n_iter"], o["n_iter"])
            assert r["final_delta"] <= 1.0
            rel = np.abs(r["fpkm"] - o["fpkm"]) / np.maximum(np.abs(o["fpkm"]), 1e-300)
            assert np.all((rel <= 1e-9) | (np.abs(r["ireadcount"] - o["ireadcount"]) <= 1e-9)), (short, rel.max())
        ix.close()
    finally:
        c.close()


@pytest.mark.parametrize("name", ["se_longk", "pe_nf21"])
def test_count_compact_wire_form(ctx, name):
    """emsar_sample_count_compact (uint16 lengths, uint16 or no fragment lengths) counts exactly what the wide form counts; a batch
    whose lengths and tid count disagree is an error."""
    from emsar_b200._lib import EmsarError
    idx, reads = _make(name)
    R0, F0, N0 = _oracle().count(idx, reads)
    ix = Index(ctx, idx)
    s = ix.sample()
    ln = np.diff(reads.read_ptr).astype(np.uint16)
    n = len(ln)
    h = n // 2
    cut = int(reads.read_ptr[h])
    if idx.nF == 1:
        s.count_compact(ln[:h], reads.read_tid[:cut], None, int(reads.read_fraglen[0]))
        s.count_compact(ln[h:], reads.read_tid[cut:], None, int(reads.read_fraglen[0]))
    else:
        s.count_compact(ln[:h], reads.read_tid[:cut], reads.read_fraglen[:h].astype(np.uint16))
        s.count_compact(ln[h:], reads.read_tid[cut:], reads.read_fraglen[h:].astype(np.uint16))
    R, F, N = s.counts()
    s.close()
    assert N == N0 and np.array_equal(F, F0) and np.array_equal(R, R0)
    s = ix.sample()
    s.count_compact(ln, reads.read_tid[:len(reads.read_tid) - 5], None, int(reads.read_fraglen[0]))
    with pytest.raises(EmsarError):
        s.counts()
    s.close(); ix.close()
