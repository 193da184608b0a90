"""Index construction on the device (csrc/build.cu: emsar_build_classes_run): `emsar-build --device 0` must write the
file the UNMODIFIED reference `emsar-build` writes - byte for byte - on the library layouts of tests/test_build_index_cpu.py (single-end
unstranded / stranded / read-length range, paired-end unstranded / stranded / RefSeq headers with a low MAX_REPEAT), also when the
paired-end candidates are processed in many partitions, and `emsar -x` (which builds on the device by default) must print the same index.
The reference binary travels as oracle/_ref/emsar-build; the committed fixtures tests/golden/build_*.rsh.gz pin it as well."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from test_build_index_cpu import CASES, GOLD, MINE, REF, ROOT, _write_sam, make_fasta

pytestmark = pytest.mark.gpu


def _run_build(args, fa, outdir, env=None, tool=MINE, device=True):
    a = [x.replace("FA", fa) if x == "FA" else x for x in args]
    a = [x if x != "OUT" else outdir for x in a]
    if tool == REF:
        a = [x for i, x in enumerate(a) if x != "-p" and (i == 0 or a[i - 1] != "-p")]      # the reference's threaded PE build loses updates
    r = subprocess.run([tool] + (["--device", "0"] if device else []) + a, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return open(os.path.join(outdir, "x.rsh"), "rb").read()


@pytest.mark.parametrize("name", list(CASES))
def test_device_build_matches_reference(built, tmp_path, name):
    fa = str(tmp_path / "t.fa")
    make_fasta(fa, refseq="refseq" in name)
    mine = _run_build(CASES[name], fa, str(tmp_path / "dev"))
    want = gzip.open(os.path.join(GOLD, f"build_{name}.rsh.gz"), "rb").read()
    assert want.count(b"\n") > 60
    assert mine == want
    if os.path.exists(REF):
        assert _run_build(CASES[name], fa, str(tmp_path / "ref"), tool=REF, device=False) == want
    assert _run_build(CASES[name], fa, str(tmp_path / "host"), device=False) == want


@pytest.mark.parametrize("name", ["pe_ns", "pe_ssrf", "pe_refseq_k4"])
def test_device_build_in_partitions(built, tmp_path, name):
    """A tiny entry buffer forces the paired-end candidates through many partitions of the mate-1 hash: same file."""
    fa = str(tmp_path / "t.fa")
    make_fasta(fa, refseq="refseq" in name)
    env = dict(os.environ, EMSAR_BUILD_CAP="1500")
    mine = _run_build(CASES[name], fa, str(tmp_path / "dev"), env=env)
    assert mine == gzip.open(os.path.join(GOLD, f"build_{name}.rsh.gz"), "rb").read()


def _big_fasta(path, seed=11, genes=260):
    rng = np.random.default_rng(seed)
    exons = ["".join(rng.choice(list("ACGT"), size=int(rng.integers(80, 300)))) for _ in range(900)]
    t = 0
    with open(path, "w") as f:
        for g in range(genes):
            pool = [exons[(5 * g + j) % len(exons)] for j in range(7)]
            for iso in range(int(rng.integers(1, 4))):
                keep = [e for e in pool if rng.random() < 0.7] or pool[:1]
                f.write(f">T{t}\n{''.join(keep)}\n")
                t += 1
    return t


@pytest.mark.parametrize("args", [["FA", "40-41", "OUT", "x"], ["-s", "ssf", "FA", "50", "OUT", "x"], ["-P", "-f", "150", "-F", "260", "FA", "40", "OUT", "x"],
                                  ["-P", "-s", "ssfr", "-f", "100", "-F", "180", "FA", "36", "OUT", "x"]])
def test_device_build_on_a_larger_transcriptome(built, tmp_path, args):
    """~500 transcripts / 400K bases, > 10^5 classes for the paired-end layouts: device = host builder = reference (where it is present)."""
    fa = str(tmp_path / "big.fa")
    _big_fasta(fa)
    dev = _run_build(["-q"] + args, fa, str(tmp_path / "dev"))
    host = _run_build(["-q", "-p", "4"] + args, fa, str(tmp_path / "host"), device=False)
    assert dev.count(b"\n") > 1000
    assert dev == host
    if os.path.exists(REF) and "-P" not in args:            # the reference's PE build of this size takes minutes
        assert _run_build(["-q"] + args, fa, str(tmp_path / "ref"), tool=REF, device=False) == dev


@pytest.mark.parametrize("pe", [False, True])
def test_emsar_dash_x_builds_on_the_device(built, tmp_path, pe):
    emsar = os.path.join(ROOT, "emsar_b200", "bin", "emsar")
    fa = str(tmp_path / "t.fa")
    tx = make_fasta(fa)
    L = 25
    _write_sam(str(tmp_path / "in.sam"), tx, L, pe, np.random.default_rng(3))
    flags = ["-q", "-R", "-S"] + (["-P", "-f", "40", "-F", "70"] if pe else [])
    outs = {}
    for tag, env in (("device", dict(os.environ)), ("host", dict(os.environ, EMSAR_BUILD_HOST="1"))):
        out = str(tmp_path / tag)
        r = subprocess.run([emsar] + flags + ["-x", fa, out, "p", str(tmp_path / "in.sam")], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[tag] = open(os.path.join(out, "p.rsh"), "rb").read()
    assert outs["device"] == outs["host"]
    assert outs["device"] == gzip.open(os.path.join(GOLD, "build_pe_ns.rsh.gz" if pe else "build_se_ns.rsh.gz"), "rb").read()


def test_device_build_api_errors(built, ctx):
    import ctypes as C
    from emsar_b200 import _lib
    L = _lib.lib()
    out = (C.c_byte * 256)()
    assert L.emsar_build_classes_run(ctx._h, None, out) != 0
    assert b"NULL" in L.emsar_cuda_last_error()
