"""Index construction from a fasta (SURVEY.md §3.4, §8 f4 host side): `emsar_b200/bin/emsar-build` must write the file the
UNMODIFIED reference `emsar-build` writes - byte for byte - for single-end (unstranded, stranded, read-length range) and
paired-end (unstranded, stranded) libraries, on a transcriptome with shared exons, paralogs, internal repeats, duplicated
transcripts, reverse-complement pairs, N runs, lower case and transcripts shorter than the read length.
The reference binary is compiled from /root/reference by oracle/Makefile; where it is absent (GPU box without oracle/_ref)
the committed fixtures tests/golden/build_*.rsh.gz (made by this very test with EMSAR_WRITE_GOLDEN=1) stand in."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINE = os.path.join(ROOT, "emsar_b200", "bin", "emsar-build")
REF = os.path.join(ROOT, "oracle", "_ref", "emsar-build")
GOLD = os.path.join(ROOT, "tests", "golden")


def make_fasta(path, seed=5, refseq=False):
    rng = np.random.default_rng(seed)
    rand = lambda n: "".join(rng.choice(list("ACGT"), size=n))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    rcs = lambda s: "".join(comp[c] for c in reversed(s))
    exons = [rand(int(rng.integers(30, 90))) for _ in range(40)]
    tx = []
    for g in range(18):                                   # gene families: isoforms share exons
        pool = [exons[(2 * g + j) % len(exons)] for j in range(4)]
        for iso in range(int(rng.integers(2, 5))):
            keep = [e for e in pool if rng.random() < 0.75] or pool[:1]
            tx.append("".join(keep))
    tx.append(tx[1][:70] + tx[1][:70] + rand(20))        # internal repeat
    tx.append(tx[4])                                     # duplicated transcript
    tx.append(rcs(tx[2]))                                # reverse complement of another transcript
    tx.append(rand(40) + "NNNNN" + rand(60) + "N" + rand(30))
    tx.append(rand(12))                                  # shorter than any read length used here
    tx.append(("ACGT" * 40))                             # low complexity: hits MAX_REPEAT
    tx.append(rand(90).lower())                          # lower case is upper-cased
    pal = rand(25)
    tx.append(pal + rcs(pal) + rand(10))                 # contains its own reverse complement
    with open(path, "w") as f:
        for i, s in enumerate(tx):
            name = f">gi|{i}|ref|TX{i:03d}.1| some text" if refseq else f">TX{i:03d} gene:G{i // 3}"
            f.write(name + "\n")
            for o in range(0, len(s), 50):               # wrapped lines
                f.write(s[o:o + 50] + "\n")
    return tx


CASES = {
    "se_ns": ["-q", "FA", "25", "OUT", "x"],
    "se_ssf": ["-q", "-s", "ssf", "FA", "25", "OUT", "x"],
    "se_range_k5": ["-q", "-k", "5", "FA", "24-27", "OUT", "x"],
    "pe_ns": ["-q", "-P", "-f", "40", "-F", "70", "FA", "25", "OUT", "x"],
    "pe_ns_threads": ["-q", "-P", "-p", "3", "-f", "40", "-F", "70", "FA", "25", "OUT", "x"],
    "pe_ssrf": ["-q", "-P", "-s", "ssrf", "-f", "30", "-F", "55", "FA", "20", "OUT", "x"],
    "pe_refseq_k4": ["-q", "-P", "-h", "R", "-k", "4", "-f", "1", "-F", "45", "FA", "22", "OUT", "x"],
}


@pytest.mark.parametrize("name", list(CASES))
def test_emsar_build_matches_reference(built, tmp_path, name):
    fa = str(tmp_path / "t.fa")
    make_fasta(fa, refseq="refseq" in name)
    args = [a.replace("FA", fa) if a == "FA" else a for a in CASES[name]]
    mine_dir = str(tmp_path / "mine")
    r = subprocess.run([MINE] + [a if a != "OUT" else mine_dir for a in args], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    mine = open(os.path.join(mine_dir, "x.rsh"), "rb").read()
    gold = os.path.join(GOLD, f"build_{name}.rsh.gz")
    if os.path.exists(REF):
        ref_dir = str(tmp_path / "ref")
        # the reference always runs single-threaded here: its threaded paired-end construction increments the singleton counts
        # without a lock (update_rshbucket_single from process_mate1_cluster_by_mate_3) and loses updates from run to run
        ref_args = [a for i, a in enumerate(args) if a != "-p" and (i == 0 or args[i - 1] != "-p")]
        r = subprocess.run([REF] + [a if a != "OUT" else ref_dir for a in ref_args], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        want = open(os.path.join(ref_dir, "x.rsh"), "rb").read()
        if os.environ.get("EMSAR_WRITE_GOLDEN"):
            with gzip.GzipFile(gold, "wb", mtime=0) as g:
                g.write(want)
        if os.path.exists(gold):
            assert gzip.open(gold, "rb").read() == want, "committed fixture is stale"
    else:
        want = gzip.open(gold, "rb").read()
    assert want.count(b"\n") > 60
    assert mine == want


def test_build_then_load_round_trip(built, tmp_path):
    from emsar_b200 import host
    fa = str(tmp_path / "t.fa")
    tx = make_fasta(fa)
    out = str(tmp_path / "o")
    assert subprocess.run([MINE, "-q", "-P", "-f", "40", "-F", "70", fa, "25", out, "x"]).returncode == 0
    r = host.Rsh(os.path.join(out, "x.rsh"))
    assert r.T == len(tx) and r.readlength == 25 and (r.frag_min, r.frag_max) == (40, 70) and r.nF == 31
    assert r.names[3] == "TX003" and r.tid("TX005") == 5
    k = np.diff(r.class_ptr)
    assert (k[:r.T] == 1).all() and (k[r.T:] >= 2).all()
    # every fragment position of an identifiable transcript is counted exactly once somewhere
    assert r.euma.sum() > 0
    r.close()


def _write_sam(path, tx, L, pe, rng):
    names = [f"TX{i:03d}" for i in range(len(tx))]
    with open(path, "w") as f:
        f.write("@HD\tVN:1.0\tSO:unsorted\n")
        for n, s in zip(names, tx):
            f.write(f"@SQ\tSN:{n}\tLN:{len(s)}\n")
        f.write("unal\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t*\n")                      # unaligned records are skipped by the sniffers
        for r in range(200):
            t = int(rng.integers(0, len(tx)))
            if len(tx[t]) < 80:
                continue
            p = int(rng.integers(0, len(tx[t]) - 70))
            rd = "A" * L
            if not pe:
                f.write(f"r{r}\t0\t{names[t]}\t{p + 1}\t255\t{L}M\t*\t0\t0\t{rd}\t*\tMD:Z:{L}\n")
            else:
                d = int(rng.integers(20, 40))
                f.write(f"r{r}\t{0x1 | 0x2 | 0x20 | 0x40}\t{names[t]}\t{p + 1}\t255\t{L}M\t=\t{p + 1 + d}\t{d + L}\t{rd}\t*\tMD:Z:{L}\n")
                f.write(f"r{r}\t{0x1 | 0x2 | 0x10 | 0x80}\t{names[t]}\t{p + 1 + d}\t255\t{L}M\t=\t{p + 1}\t{-(d + L)}\t{rd}\t*\tMD:Z:{L}\n")


@pytest.mark.parametrize("pe", [False, True])
def test_emsar_dash_x_builds_the_same_index(built, tmp_path, pe):
    """`emsar -R -x t.fa ...` learns the read length from the alignment file and prints the index `emsar-build` prints
    (SURVEY.md §3.4). Without a GPU the run stops right after that, at emsar_cuda_open: there is no CPU path for the rest."""
    emsar = os.path.join(ROOT, "emsar_b200", "bin", "emsar")
    fa = str(tmp_path / "t.fa")
    tx = make_fasta(fa)
    L = 25
    _write_sam(str(tmp_path / "in.sam"), tx, L, pe, np.random.default_rng(3))
    out = str(tmp_path / "out")
    flags = ["-q", "-R", "-S"] + (["-P", "-f", "40", "-F", "70"] if pe else [])
    import torch
    # with a GPU `emsar -x` constructs the classes on the device (tests/test_build_index_gpu.py); EMSAR_BUILD_HOST=1 keeps them on the host
    env = dict(os.environ) if torch.cuda.is_available() else dict(os.environ, EMSAR_BUILD_HOST="1")
    r = subprocess.run([emsar] + flags + ["-x", fa, out, "p", str(tmp_path / "in.sam")], capture_output=True, text=True, env=env)
    if not torch.cuda.is_available():
        assert r.returncode != 0 and "no CPU" in (r.stdout + r.stderr)
    else:
        assert r.returncode == 0, r.stdout + r.stderr
    mine = open(os.path.join(out, "p.rsh"), "rb").read()
    bdir = str(tmp_path / "b")
    bflags = ["-q"] + (["-P", "-f", "40", "-F", "70"] if pe else [])
    tool = REF if os.path.exists(REF) else MINE
    assert subprocess.run([tool] + bflags + [fa, str(L), bdir, "x"], capture_output=True).returncode == 0
    assert mine == open(os.path.join(bdir, "x.rsh"), "rb").read()


def test_emsar_dash_x_se_read_length_range_from_bowtie(built, tmp_path):
    """SE: -x scans the whole alignment file for the shortest and the longest read (read_bowtie_get_readlengths_se :260-288)
    and builds one EUMA column per length."""
    emsar = os.path.join(ROOT, "emsar_b200", "bin", "emsar")
    fa = str(tmp_path / "t.fa")
    tx = make_fasta(fa)
    rng = np.random.default_rng(4)
    with open(tmp_path / "in.bowtie", "w") as f:
        for r in range(150):
            t = int(rng.integers(0, len(tx)))
            L = int(rng.integers(24, 28)) if r not in (0, 1) else (24, 27)[r]      # both ends of the range are present
            if len(tx[t]) < 40:
                continue
            f.write(f"r{r}\t+\tTX{t:03d}\t{int(rng.integers(0, len(tx[t]) - 30))}\t{'A' * L}\t{'I' * L}\t0\t\n")
    out = str(tmp_path / "out")
    import torch
    env = dict(os.environ) if torch.cuda.is_available() else dict(os.environ, EMSAR_BUILD_HOST="1")
    subprocess.run([emsar, "-q", "-R", "-x", fa, out, "p", str(tmp_path / "in.bowtie")], capture_output=True, text=True, env=env)
    mine = open(os.path.join(out, "p.rsh"), "rb").read()
    assert mine.startswith(b"#") and b",24,27,-1\n" in mine.split(b"\n")[0] + b"\n"
    bdir = str(tmp_path / "b")
    tool = REF if os.path.exists(REF) else MINE
    assert subprocess.run([tool, "-q", fa, "24-27", bdir, "x"], capture_output=True).returncode == 0
    assert mine == open(os.path.join(bdir, "x.rsh"), "rb").read()


def test_fasta_errors(built, tmp_path):
    bad = tmp_path / "bad.fa"
    bad.write_text("ACGT\n>t\nACGT\n")                       # does not start with '>'
    r = subprocess.run([MINE, "-q", str(bad), "25", str(tmp_path / "o"), "x"], capture_output=True, text=True)
    assert r.returncode != 0 and "wrong fasta file format" in r.stderr
    r = subprocess.run([MINE, "-q", str(tmp_path / "none.fa"), "25", str(tmp_path / "o"), "x"], capture_output=True, text=True)
    assert r.returncode != 0 and "can't open fasta file" in r.stderr
    r = subprocess.run([MINE, "-q", "-s", "ssfr", str(bad), "25", str(tmp_path / "o"), "x"], capture_output=True, text=True)
    assert r.returncode != 0 and "invalid strand type" in r.stderr      # a PE strand type without -P


def test_threaded_single_end_build_on_a_larger_transcriptome(built, tmp_path):
    """> 200K substrings: the single-end construction partitions by hash and runs on several threads; same bytes as with one
    thread and as the reference."""
    rng = np.random.default_rng(11)
    exons = ["".join(rng.choice(list("ACGT"), size=int(rng.integers(80, 300)))) for _ in range(900)]
    fa = str(tmp_path / "big.fa")
    with open(fa, "w") as f:
        t = 0
        for g in range(260):
            pool = [exons[(5 * g + j) % len(exons)] for j in range(7)]
            for iso in range(int(rng.integers(1, 4))):
                keep = [e for e in pool if rng.random() < 0.7] or pool[:1]
                f.write(f">T{t}\n{''.join(keep)}\n")
                t += 1
    outs = {}
    for tag, tool, extra in (("p4", MINE, ["-p", "4"]), ("p1", MINE, []), ("ref", REF, [])):
        if not os.path.exists(tool):
            continue
        d = str(tmp_path / tag)
        assert subprocess.run([tool, "-q"] + extra + [fa, "40-41", d, "x"], capture_output=True).returncode == 0
        outs[tag] = open(os.path.join(d, "x.rsh"), "rb").read()
    assert outs["p4"].count(b"\n") > 1000
    assert len(set(outs.values())) == 1, {k: len(v) for k, v in outs.items()}
