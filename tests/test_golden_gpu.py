"""GPU tests against outputs of the UNMODIFIED reference binary: the `emsar` command-line program of this repo
(host C + libemsar_cuda) is run on the golden fixtures with the reference's own flags and its output files are
compared with the reference's files (tests/golden)."""
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMSAR = os.path.join(ROOT, "emsar_b200", "bin", "emsar")


def _run_cli(fx, tmp_path, extra=()):
    rsh = gu.materialize(fx["rsh"], tmp_path)
    aln = gu.materialize(fx["aln"], tmp_path)
    out = os.path.join(str(tmp_path), "out")
    cmd = [EMSAR, "-q", "-g", "-k", str(fx["k"]), "-s", fx["strand"]]
    if fx["pe"]:
        cmd.append("-P")
    if fx["fmt"] == "sam":
        cmd.append("-S")
    elif fx["fmt"] == "bam":
        cmd.append("-B")
    cmd += list(extra) + ["-I", rsh, out, "p", aln]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return out


@pytest.mark.parametrize("name", list(gu.FIXTURES))
def test_cli_outputs_match_reference_files(built, name, tmp_path):
    fx = gu.FIXTURES[name]
    out = _run_cli(fx, tmp_path)
    seg_ref, fl_ref = gu.read_segments(fx["out"]), gu.read_fraglength(fx["out"])
    # .fraglength_effect: identical text (integer counts, %lg of bit-identical Wf)
    assert gu.parse_out_file(os.path.join(out, "p.0.fraglength_effect")) == fl_ref["text"]
    # .segments: ids, set ids, transcript lists, names, eff.length (%lf) and Readcount identical; expected within tolerance
    seg = gu.parse_out_file(os.path.join(out, "p.0.segments"))
    assert [r[:6] for r in seg] == [r[:6] for r in seg_ref["text"]]
    N = int(fl_ref["counts"].sum())
    if not gu.has_fpkm(fx["out"]):
        # no .fpkm kept for this fixture: expected counts within the north-star 1e-6 relative / 1e-3 reads (+ the two files' 6 decimals)
        ex, ex_ref = np.array([float(r[6]) for r in seg]), seg_ref["expected"]
        assert (np.abs(ex - ex_ref) <= np.maximum(1e-6 * np.abs(ex_ref), 1e-3) + 2 * gu.PRINT_EPS).all()
        return
    g = gu.read_fpkm(fx["out"])
    mine = gu.parse_out_file(os.path.join(out, "p.0.fpkm"))
    assert [r[0] for r in mine] == g["names"]
    assert [r[3] for r in mine] == [f"{e:f}" for e in g["efflen"]]          # eff.length: identical text
    fp = np.array([float(r[1]) for r in mine])
    tol = gu.fpkm_tolerance(g, g["efflen"], max(N, 1), fx["rounds"], from_files=True)     # max(1e-6 rel, 1e-3 reads, 6 sd) + the files' 6 decimals
    ident = seg_ref["adjEUMA"][:len(fp)] > 0
    assert (np.abs(fp - g["fpkm"])[ident] <= tol[ident]).all()
    ir = np.array([float(r[4]) for r in mine])
    tol_ir = gu.ireadcount_tolerance(tol, g["efflen"], max(N, 1), from_files=True)
    assert (np.abs(ir - g["ireadcount"])[ident] <= tol_ir[ident]).all()
    # per-class expected counts (unique at the optimum even where FPKM is not): the members' FPKM tolerances carried through the sum
    cp = np.concatenate([[0], np.cumsum([len(r[2].split(",")) for r in seg_ref["text"]])])
    ct = np.array([int(t[1:]) for r in seg_ref["text"] for t in r[2].split(",")])
    tol_ex = gu.expected_tolerance(np.where(ident, tol, 0.0), cp, ct, seg_ref["adjEUMA"], max(N, 1), from_files=True)
    ex = np.array([float(r[6]) for r in seg])
    assert (np.abs(ex - seg_ref["expected"]) <= tol_ex).all(), float((np.abs(ex - seg_ref["expected"]) - tol_ex).max())
    if ident.all():
        tpm = np.array([float(r[6]) for r in mine])
        tol_tpm = tol * 1e6 / g["fpkm"].sum() + np.abs(g["tpm"]) * tol[ident].sum() / g["fpkm"].sum() + 2 * gu.PRINT_EPS   # TPM = FPKM * 1e6 / sum FPKM (:3207)
        assert (np.abs(tpm - g["tpm"]) <= tol_tpm).all()


def test_cli_multisample_and_print_rsh(built, tmp_path):
    """-M: every file of the list is an independent sample (emsar_main.c:380-488); -R rewrites the index."""
    fx = gu.FIXTURES["se"]
    rsh = gu.materialize(fx["rsh"], tmp_path)
    aln = gu.materialize(fx["aln"], tmp_path)
    half = os.path.join(str(tmp_path), "half.bowtie")
    lines = open(aln).read().splitlines(True)
    open(half, "w").writelines(lines[:len(lines) // 2 // 2 * 2])
    lst = os.path.join(str(tmp_path), "list.txt")
    open(lst, "w").write(aln + "\n" + half + "\n" + aln + "\n")
    out = os.path.join(str(tmp_path), "outM")
    r = subprocess.run([EMSAR, "-q", "-g", "-M", "-R", "-I", rsh, out, "p", lst], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a = open(os.path.join(out, "p.0.fpkm")).read()
    assert a == open(os.path.join(out, "p.2.fpkm")).read()          # deterministic: same sample twice, same bytes
    assert a != open(os.path.join(out, "p.1.fpkm")).read()
    assert gu.parse_out_file(os.path.join(out, "p.0.fraglength_effect")) == gu.read_fraglength("se")["text"]
    assert open(os.path.join(out, "p.rsh")).read() == open(rsh).read()


def test_cli_errors_like_reference(built, tmp_path):
    fx = gu.FIXTURES["se"]
    rsh = gu.materialize(fx["rsh"], tmp_path)
    bad = os.path.join(str(tmp_path), "bad.bowtie")
    open(bad, "w").write("r1\t+\tNOT_A_TRANSCRIPT\t0\tA\tA\t0\t\n")
    r = subprocess.run([EMSAR, "-q", "-I", rsh, os.path.join(str(tmp_path), "o"), "p", bad], capture_output=True, text=True)
    assert r.returncode == 1 and "unexisting tid" in r.stderr
    r = subprocess.run([EMSAR, "-q", "-k", "5000", "-I", rsh, os.path.join(str(tmp_path), "o"), "p", bad], capture_output=True, text=True)
    assert r.returncode == 1 and "1024" in r.stderr
    r = subprocess.run([EMSAR, "-q", "-I", os.path.join(str(tmp_path), "none.rsh"), os.path.join(str(tmp_path), "o"), "p", bad], capture_output=True, text=True)
    assert r.returncode == 1


def test_eumacut_fixture_through_cuda(ctx, built, tmp_path):
    from emsar_b200 import host
    from emsar_b200.api import Index
    rsh = host.Rsh(gu.materialize("eumacut.in.rsh", tmp_path))
    reads, _ = host.read_alignments(rsh, gu.materialize("eumacut.in.bowtie", tmp_path))
    ref = np.load(os.path.join(gu.GOLD, "eumacut.ref.npz"))
    ix = Index(ctx, rsh)
    assert ix.info()["max_set_tids"] > 5000
    s = ix.sample()
    s.count(reads.read_ptr, reads.read_tid, reads.read_fraglen)
    R, F, N = s.counts()
    r = s.solve()
    adj, ex, cs = s.segments()
    s.close(); ix.close(); rsh.close()
    assert np.array_equal(R, ref["ReadCount"])
    assert r["eumacut"] == 8.0
    assert np.array_equal(cs, ref["set_id"])
    assert [f"{a:f}" for a in adj] == [f"{a:f}" for a in ref["adjEUMA"]]


def test_live_reference_binary_if_present(built, tmp_path):
    """When oracle/_ref/emsar travelled to this box: run it and this repo's emsar on a fresh seeded fixture."""
    from emsar_b200 import synth
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "emsar")
    if not os.path.exists(ref_bin):
        pytest.skip("reference binary not present on this box")
    idx = synth.make_index(T=400, n_multi=2500, kmax=15, seed=77, module_cap=40)
    reads = synth.make_reads(idx, 15000, seed=77)
    d = str(tmp_path)
    synth.write_rsh(idx, d + "/x.rsh"); synth.write_bowtie_se(idx, reads, d + "/x.bowtie")
    a = subprocess.run([ref_bin, "-q", "-g", "-n", "6", "-I", d + "/x.rsh", d + "/ref", "p", d + "/x.bowtie"], capture_output=True, text=True)
    b = subprocess.run([EMSAR, "-q", "-g", "-I", d + "/x.rsh", d + "/mine", "p", d + "/x.bowtie"], capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert open(d + "/ref/p.0.fraglength_effect").read() == open(d + "/mine/p.0.fraglength_effect").read()
    sr, sm = gu.parse_out_file(d + "/ref/p.0.segments"), gu.parse_out_file(d + "/mine/p.0.segments")
    assert [r[:6] for r in sr] == [r[:6] for r in sm]
    fr, fm = gu.parse_out_file(d + "/ref/p.0.fpkm"), gu.parse_out_file(d + "/mine/p.0.fpkm")
    g = dict(fpkm=np.array([float(r[1]) for r in fr]), sd=np.array([float(r[2]) for r in fr]))
    eff = np.array([float(r[3]) for r in fr])
    tol = gu.fpkm_tolerance(g, eff, 15000, 6)
    assert (np.abs(np.array([float(r[1]) for r in fm]) - g["fpkm"]) <= tol).all()


def test_cli_restart_rounds_sd_column(built, tmp_path):
    """-n R: R - 1 restart rounds from seeded random starts. Where the optimum is unique every round lands on it (sd.of.FPKM ~ 0, same
    FPKM as the single run); the duplicated transcripts of the `built` fixture (no private k-mer) are what the data cannot tell apart:
    their sd is large while their SUM is the same in every round."""
    fx = gu.FIXTURES["built"]
    rsh, aln = gu.materialize(fx["rsh"], tmp_path), gu.materialize(fx["aln"], tmp_path)
    outs = {}
    for tag, extra in (("one", []), ("n4", ["-n", "4"])):
        out = os.path.join(str(tmp_path), tag)
        r = subprocess.run([EMSAR, "-q", "-g", "-S", "-s", fx["strand"]] + extra + ["-I", rsh, out, "p", aln], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[tag] = gu.parse_out_file(os.path.join(out, "p.0.fpkm"))
    one, n4 = outs["one"], outs["n4"]
    seg = gu.read_segments(fx["out"])
    f1, f4 = np.array([float(r[1]) for r in one]), np.array([float(r[1]) for r in n4])
    sd1, sd4 = np.array([float(r[2]) for r in one]), np.array([float(r[2]) for r in n4])
    ident = seg["adjEUMA"][:len(f1)] > 0
    assert (sd1 == 0).all()
    # every round stops within the EM tolerance of the same optimum: the spread is orders of magnitude below the estimates
    assert np.allclose(f1[ident], f4[ident], rtol=1e-3, atol=1e-3) and (sd4[ident] <= 1e-3 * np.maximum(f4[ident], 1.0)).all()
    amb = ~ident
    assert amb.sum() >= 2 and abs(f1[amb].sum() - f4[amb].sum()) <= 1e-3 * max(f1[amb].sum(), 1.0) + 1e-3
    assert sd4[amb].max() > 100 * max(sd4[ident].max(), 1e-9)            # the split between indistinguishable transcripts depends on the start
    assert open(os.path.join(str(tmp_path), "one", "p.0.segments")).read() == open(os.path.join(str(tmp_path), "n4", "p.0.segments")).read()


def test_cli_complete_index_image(built, tmp_path):
    """with EMSAR_RSH_CACHE=1 the first run writes <rsh>.pack including what the device library derived (transpose,
    locality order, reachable classes, set statistics); the second run creates the index from the image without deriving anything and
    writes the same files byte for byte."""
    fx = gu.FIXTURES["config1"]
    rsh, aln = gu.materialize(fx["rsh"], tmp_path), gu.materialize(fx["aln"], tmp_path)
    env = dict(os.environ, EMSAR_RSH_CACHE="1")
    outs = []
    for tag in ("first", "second"):
        out = os.path.join(str(tmp_path), tag)
        r = subprocess.run([EMSAR, "-g", "-P", "-S", "-I", rsh, out, "p", aln], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append((out, r.stdout))
    assert "packed index image written" in outs[0][1] and os.path.exists(rsh + ".pack")
    assert "rsh index taken from its packed image" in outs[1][1] and "nothing derived at load" in outs[1][1]
    for ext in ("fpkm", "segments", "fraglength_effect"):
        assert open(os.path.join(outs[0][0], f"p.0.{ext}")).read() == open(os.path.join(outs[1][0], f"p.0.{ext}")).read(), ext
