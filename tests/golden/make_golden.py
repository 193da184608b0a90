#!/usr/bin/env python
"""Generates tests/golden/*: fixture inputs plus the outputs of the UNMODIFIED reference (oracle/_ref/emsar,
oracle/_ref/emsar-build, built from /root/reference/src by oracle/Makefile) run on them.

The reference ships no tests or golden vectors for this path, so these files are what pins the
oracle and the CUDA path to the reference's behaviour. Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Everything is seeded; the only non-deterministic part is the reference's estimator (srand(time)), which is why the
.fpkm files are produced with -n 8 and compared under the tolerance policy of SURVEY.md §8(c).
"""
import gzip
import os
import shutil
import struct
import subprocess
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from emsar_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "emsar")
REFB = os.path.join(ROOT, "oracle", "_ref", "emsar-build")


def run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("%s\n%s\n%s" % (" ".join(cmd), r.stdout[-2000:], r.stderr[-2000:]))
    return r


def gz(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


def keep(tmp, name, files):
    for f in files:
        gz(os.path.join(tmp, f), os.path.join(HERE, f"{name}.{os.path.basename(f)}.gz"))


from emsar_b200.synth import sam_to_bam  # noqa: E402  (the minimal BAM writer lives next to the other fixture writers)


def fixture_se(tmp):
    idx = synth.make_index(T=300, n_multi=1500, kmax=12, seed=1, module_cap=40)
    reads = synth.make_reads(idx, 8000, seed=1)
    synth.write_rsh(idx, f"{tmp}/in.rsh")
    synth.write_bowtie_se(idx, reads, f"{tmp}/in.bowtie")
    run([REF, "-q", "-g", "-n", "8", "-I", f"{tmp}/in.rsh", f"{tmp}/out", "p", f"{tmp}/in.bowtie"])
    keep(tmp, "se", ["in.rsh", "in.bowtie", "out/p.0.fpkm", "out/p.0.segments", "out/p.0.fraglength_effect"])


def fixture_pe(tmp):
    idx = synth.make_index(T=200, n_multi=900, kmax=8, seed=3, module_cap=30, nF=21, frag_min=40, readlength=25)
    reads = synth.make_reads(idx, 5000, seed=3)
    synth.write_rsh(idx, f"{tmp}/in.rsh")
    synth.write_sam_pe(idx, reads, f"{tmp}/in.sam")
    run([REF, "-q", "-g", "-n", "8", "-P", "-S", "-I", f"{tmp}/in.rsh", f"{tmp}/out", "p", f"{tmp}/in.sam"])
    keep(tmp, "pe", ["in.rsh", "in.sam", "out/p.0.fpkm", "out/p.0.segments", "out/p.0.fraglength_effect"])
    # the same alignments as BAM, stranded ssfr: pins the BGZF/BAM decoder and the strand filter
    sam_to_bam(f"{tmp}/in.sam", f"{tmp}/in.bam")
    run([REF, "-q", "-g", "-n", "2", "-P", "-B", "-s", "ssfr", "-I", f"{tmp}/in.rsh", f"{tmp}/outb", "p", f"{tmp}/in.bam"])
    shutil.copy(f"{tmp}/in.bam", os.path.join(HERE, "pe.in.bam"))
    keep(tmp, "pe_bam_ssfr", ["outb/p.0.segments", "outb/p.0.fraglength_effect"])


def fixture_crafted(tmp):
    """Hand-written read groups that pin SURVEY.md §3.2: duplicate removal before the mismatch test, best-mm
    filtering, the <= MAX_REPEAT gate (-k 3), an unmatched multiset (counted in N only), {t,t} != {t}, the
    fragment-length window, the strand filter, unsorted tids, a singleton without node."""
    names = ["tA", "tB", "tC", "tD", "tE"]
    with open(f"{tmp}/in.rsh", "w") as f:
        f.write("#4,3,2,3,-1\n")
        for i, n in enumerate(names):
            f.write(f"@{i}\t{n}\n")
        f.write("cid\tno.tids\tfirst.tid\tother.tids\tsegment.length\n")
        f.write("0\t1\t0\t\t10,4,\n1\t1\t1\t\t20,5,\n2\t1\t2\t\t30,6,\n3\t1\t3\t\t\t\n4\t1\t4\t\t7,0,\n")
        f.write("5\t2\t0\t0,\t3,1,\n6\t2\t0\t1,\t5,2,\n7\t2\t1\t2,\t6,2,\n8\t3\t0\t1,2,\t2,1,\n")
    L = []

    def aln(rid, strand, t, pos, seqlen, mm=""):
        s = "A" * seqlen
        L.append(f"{rid}\t{strand}\t{names[t]}\t{pos}\t{s}\t{s}\t0\t{mm}")
    aln("r1", "+", 0, 5, 2)                                   # singleton tA
    aln("r2", "+", 1, 5, 2); aln("r2", "+", 0, 9, 2)          # {tA,tB}, tids unsorted
    aln("r3", "+", 0, 1, 2); aln("r3", "+", 0, 7, 2)          # {tA,tA}: internal repeat class 5, not the singleton
    aln("r4", "+", 2, 1, 2, "1:A>C"); aln("r4", "+", 1, 1, 2) # best-mm filter keeps only tB
    aln("r5", "+", 1, 3, 2); aln("r5", "+", 1, 3, 2)          # exact duplicate removed -> singleton tB
    aln("r6", "+", 0, 1, 2); aln("r6", "+", 1, 1, 2); aln("r6", "+", 2, 1, 2); aln("r6", "+", 4, 1, 2)  # 4 > -k 3: dropped entirely
    aln("r7", "+", 0, 1, 2); aln("r7", "+", 2, 1, 2)          # {tA,tC}: no such class -> N and Wf only
    aln("r8", "+", 3, 1, 2)                                   # singleton without node: N only
    aln("r9", "+", 1, 1, 5)                                   # fragment length 5 outside [2,3]: ignored entirely
    aln("r10", "-", 2, 1, 3); aln("r10", "+", 1, 1, 3)        # -s ssf drops the '-' alignment -> singleton tB, length 3
    aln("r11", "+", 2, 4, 3); aln("r11", "+", 0, 4, 3); aln("r11", "+", 1, 4, 3)   # {tA,tB,tC}
    aln("r12", "+", 1, 2, 2, "1:A>C,3:G>T"); aln("r12", "+", 2, 2, 2, "0:A>C,1:C>G")  # equal mm (2): both kept -> {tB,tC}
    aln("r13", "+", 0, 1, 2, "1:A>C"); aln("r13", "+", 0, 1, 2)  # duplicate scan precedes the mm test: 2nd dropped, mm stays 1
    aln("r13", "+", 1, 1, 2)                                  # mm 0 < 1: list reset -> singleton tB
    open(f"{tmp}/in.bowtie", "w").write("\n".join(L) + "\n")
    run([REF, "-q", "-g", "-n", "2", "-k", "3", "-s", "ssf", "-I", f"{tmp}/in.rsh", f"{tmp}/out", "p", f"{tmp}/in.bowtie"])
    keep(tmp, "crafted", ["in.rsh", "in.bowtie", "out/p.0.fpkm", "out/p.0.segments", "out/p.0.fraglength_effect"])


def fixture_bowtie_pe(tmp):
    """bowtie-format PE with the mate-order quirk (emsar_functions.c:652): pins which orientation survives -s ssfr."""
    names = ["tA", "tB"]
    with open(f"{tmp}/in.rsh", "w") as f:
        f.write("#1,2,40,60,25\n@0\ttA\n@1\ttB\ncid\tno.tids\tfirst.tid\tother.tids\tsegment.length\n")
        e = "".join(f"{21 - i}," for i in range(21))
        f.write(f"0\t1\t0\t\t{e}\n1\t1\t1\t\t{e}\n2\t2\t0\t1,\t{e}\n")
    s = "A" * 25
    L = []

    def pair(rid, t, s1, p1, s2, p2):
        L.append(f"{rid}/1\t{s1}\t{names[t]}\t{p1}\t{s}\t{s}\t0\t")
        L.append(f"{rid}/2\t{s2}\t{names[t]}\t{p2}\t{s}\t{s}\t0\t")
    pair("q1", 0, "+", 10, "-", 30)     # /1 forward at 10, /2 reverse at 30
    pair("q2", 0, "-", 60, "+", 35)     # /1 reverse at 60, /2 forward at 35
    pair("q3", 1, "+", 5, "-", 25); pair("q3", 0, "+", 7, "-", 27)
    pair("q4", 1, "-", 50, "+", 20); pair("q4", 0, "-", 52, "+", 22)
    open(f"{tmp}/in.bowtie", "w").write("\n".join(L) + "\n")
    for st in ("ns", "ssfr", "ssrf"):
        run([REF, "-q", "-g", "-n", "2", "-P", "-s", st, "-I", f"{tmp}/in.rsh", f"{tmp}/out_{st}", "p", f"{tmp}/in.bowtie"])
        keep(tmp, f"bowtie_pe_{st}", [f"out_{st}/p.0.segments", f"out_{st}/p.0.fraglength_effect"])
    keep(tmp, "bowtie_pe", ["in.rsh", "in.bowtie"])


def fixture_built(tmp):
    """A real index: emsar-build on a generated fasta (gene families sharing exons, an internal repeat, a duplicated
    transcript), SE stranded L=30; reads are exact L-mers, aligned by dictionary lookup, written as SAM."""
    rng = np.random.default_rng(5)
    L = 30
    exons = ["".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(60, 200, size=60)]
    tx = []
    for g in range(12):
        pool = list(rng.choice(60, size=6, replace=False))
        for _ in range(int(rng.integers(2, 5))):
            ks = sorted(rng.choice(6, size=int(rng.integers(2, 5)), replace=False))
            tx.append("".join(exons[pool[k]] for k in ks))
    tx.append(tx[0][:80] + tx[0][:80] + tx[1][:50])      # internal repeat
    tx.append(tx[3])                                    # duplicated transcript (non-identifiable pair)
    names = [f"TX{i:03d}" for i in range(len(tx))]
    with open(f"{tmp}/t.fa", "w") as f:
        for n, s in zip(names, tx):
            f.write(f">{n}\n{s}\n")
    run([REFB, "-q", "-s", "ssf", f"{tmp}/t.fa", str(L), tmp, "built"])
    theta = rng.lognormal(0, 1.5, size=len(tx))
    w = np.array([max(len(s) - L + 1, 0) for s in tx]) * theta
    lut = {}
    for t, s in enumerate(tx):
        for p in range(len(s) - L + 1):
            lut.setdefault(s[p:p + L], []).append((t, p))
    n_reads = 6000
    src = rng.choice(len(tx), size=n_reads, p=w / w.sum())
    with open(f"{tmp}/in.sam", "w") as f:
        f.write("@HD\tVN:1.0\tSO:unsorted\n")
        for n, s in zip(names, tx):
            f.write(f"@SQ\tSN:{n}\tLN:{len(s)}\n")
        for r, t in enumerate(src):
            p = int(rng.integers(0, len(tx[t]) - L + 1))
            read = tx[t][p:p + L]
            for (t2, p2) in lut[read]:
                f.write(f"r{r}\t0\t{names[t2]}\t{p2 + 1}\t255\t{L}M\t*\t0\t0\t{read}\t*\tMD:Z:{L}\n")
    run([REF, "-q", "-g", "-n", "8", "-S", "-s", "ssf", "-I", f"{tmp}/built.rsh", f"{tmp}/out", "p", f"{tmp}/in.sam"])
    shutil.copy(f"{tmp}/built.rsh", f"{tmp}/in.rsh")
    keep(tmp, "built", ["in.rsh", "in.sam", "out/p.0.fpkm", "out/p.0.segments", "out/p.0.fraglength_effect"])


def fixture_eumacut(tmp):
    """One module of 5400 transcripts (> MAX_NTID_PER_SID = 5000): the reference raises EUMAcut by 2 until the sets
    fit (emsar_main.c:411-425). Only the set ids and adjEUMA are kept (-i 1 -l 1 keeps the MLE short)."""
    rng = np.random.default_rng(9)
    T = 5400
    pairs = np.stack([np.arange(T - 1), np.arange(1, T)], axis=1)                     # a chain: one connected module
    base = rng.integers(0, T - 12, size=9000)
    tri = np.unique(np.sort(base[:, None] + rng.integers(0, 12, size=(9000, 3)), axis=1), axis=0)
    far = np.unique(np.sort(np.stack([rng.integers(0, T, 400), rng.integers(0, T, 400)], axis=1), axis=1), axis=0)
    far = far[(far[:, 1] - far[:, 0]) > 1]                                            # long-range links, not chain pairs
    two = np.unique(np.concatenate([pairs, far]), axis=0)
    card = np.concatenate([np.full(len(two), 2), np.full(len(tri), 3)])
    C = T + len(card)
    class_ptr = np.zeros(C + 1, dtype=np.int64)
    class_ptr[1:T + 1] = np.arange(1, T + 1)
    class_ptr[T + 1:] = T + np.cumsum(card)
    class_tid = np.concatenate([np.arange(T), two.ravel(), tri.ravel()]).astype(np.int32)
    euma = np.concatenate([rng.integers(50, 900, size=T), rng.integers(1, 12, size=len(card))]).astype(np.int32)[:, None]
    idx = synth.SynthIndex(T=T, names=[f"T{t:05d}" for t in range(T)], class_ptr=class_ptr, class_tid=class_tid, euma=euma,
                           has_node=np.ones(C, dtype=np.uint8), min_fraglength=1, max_fraglength=1, readlength=-1, max_t_size=3)
    reads = synth.make_reads(idx, 3000, seed=9)
    synth.write_rsh(idx, f"{tmp}/in.rsh")
    synth.write_bowtie_se(idx, reads, f"{tmp}/in.bowtie")
    r = run([REF, "-g", "-n", "2", "-i", "1", "-l", "1", "-I", f"{tmp}/in.rsh", f"{tmp}/out", "p", f"{tmp}/in.bowtie"])
    cuts = [l for l in r.stdout.splitlines() if "EUMAcut is readjusted" in l]
    seg = [l.rstrip("\n").split("\t") for l in open(f"{tmp}/out/p.0.segments")][1:]
    cs = np.array([int(s[1][1:]) for s in seg], dtype=np.int32)
    adj = np.array([float(s[4]) for s in seg])
    R = np.array([int(s[5]) for s in seg], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "eumacut.ref.npz"), set_id=cs, adjEUMA=adj, ReadCount=R, n_cut_messages=len(cuts),
                        last_message=cuts[-1] if cuts else "")
    keep(tmp, "eumacut", ["in.rsh", "in.bowtie"])


def fixture_bigmod(tmp):
    """Two sequence-sharing sets of ~700 transcripts each (gene families tied together by classes across a paralog family): the size at
    which the reference's randomized search is slow (minutes per round) and its rounds visibly disagree, so the 6 sd term of the
    tolerance policy is exercised. -n 8 rounds, two sets on two threads."""
    idx = synth.make_index_v2(T=1400, n_multi=9000, alpha=2.0, kmax=24, seed=11, module_cap=700, p_cross=0.25, scatter=True)
    idx.names = [f"T{t:05d}" for t in range(idx.T)]
    reads = synth.make_reads(idx, 60000, seed=11)
    synth.write_rsh(idx, f"{tmp}/in.rsh")
    synth.write_bowtie_se(idx, reads, f"{tmp}/in.bowtie")
    run([REF, "-q", "-g", "-p", "2", "-n", "8", "-I", f"{tmp}/in.rsh", f"{tmp}/out", "p", f"{tmp}/in.bowtie"])
    keep(tmp, "bigmod", ["in.rsh", "in.bowtie", "out/p.0.fpkm", "out/p.0.segments", "out/p.0.fraglength_effect"])


def fixture_config1(tmp):
    """Substitute of BASELINE.json configs[0] (the bundled Vicugna PE sample is missing from the reference checkout, .MISSING_LARGE_BLOBS):
    a PE index built by the reference's own emsar-build -P (L = 40, fragments 60..90) from a generated 2K-transcript fasta (gene families
    sharing exons), 100K simulated PE fragments written as SAM with their true multi-mapping alignments, quantified by the reference."""
    rng = np.random.default_rng(21)
    L, fmin, fmax = 40, 60, 90
    exons = ["".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(80, 260, size=2600)]
    tx = []
    while len(tx) < 2000:
        pool = list(rng.choice(len(exons), size=8, replace=False))
        for _ in range(int(rng.integers(2, 7))):
            ks = sorted(rng.choice(8, size=int(rng.integers(2, 6)), replace=False))
            tx.append("".join(exons[pool[k]] for k in ks))
    tx = tx[:2000]
    names = [f"TX{i:04d}" for i in range(len(tx))]
    with open(f"{tmp}/tx.fa", "w") as f:
        for n, s_ in zip(names, tx):
            f.write(f">{n}\n{s_}\n")
    run([REFB, "-p", "4", "-P", "-f", str(fmin), "-F", str(fmax), f"{tmp}/tx.fa", str(L), f"{tmp}/idx", "c1"])
    rsh = [p for p in os.listdir(f"{tmp}/idx") if p.endswith(".rsh")][0]
    shutil.copy(f"{tmp}/idx/{rsh}", f"{tmp}/in.rsh")
    # fragments: position + length on a transcript drawn by expression; every (mate1 L-mer, mate2 L-mer, fragment length) occurrence in
    # any transcript is an alignment of the pair (dictionary of fragment end pairs)
    theta = rng.lognormal(0, 1.5, size=len(tx)) * (rng.random(len(tx)) > 0.2)
    w = theta * np.array([max(len(s_) - fmin + 1, 0) for s_ in tx])
    pick = rng.choice(len(tx), size=100000, p=w / w.sum())
    occ = {}
    for t, s_ in enumerate(tx):
        for fl in range(fmin, fmax + 1):
            for p0 in range(0, len(s_) - fl + 1):
                occ.setdefault((s_[p0:p0 + L], s_[p0 + fl - L:p0 + fl], fl), []).append((t, p0))
    seq = "A" * L
    with open(f"{tmp}/in.sam", "w") as f:
        f.write("@HD\tVN:1.0\tSO:unsorted\n")
        for n, s_ in zip(names, tx):
            f.write(f"@SQ\tSN:{n}\tLN:{len(s_)}\n")
        for r, t in enumerate(pick):
            s_ = tx[t]
            fl = int(rng.integers(fmin, min(fmax, len(s_)) + 1))
            p0 = int(rng.integers(0, len(s_) - fl + 1))
            for (tt, pp) in occ[(s_[p0:p0 + L], s_[p0 + fl - L:p0 + fl], fl)]:
                p1, p2 = pp + 1, pp + fl - L + 1
                f.write(f"r{r}\t{0x1 | 0x2 | 0x20 | 0x40}\t{names[tt]}\t{p1}\t255\t{L}M\t=\t{p2}\t{fl}\t{seq}\t{seq}\tMD:Z:{L}\n")
                f.write(f"r{r}\t{0x1 | 0x2 | 0x10 | 0x80}\t{names[tt]}\t{p2}\t255\t{L}M\t=\t{p1}\t{-fl}\t{seq}\t{seq}\tMD:Z:{L}\n")
    run([REF, "-q", "-g", "-p", "4", "-n", "8", "-P", "-S", "-I", f"{tmp}/in.rsh", f"{tmp}/out", "p", f"{tmp}/in.sam"])
    keep(tmp, "config1", ["in.rsh", "out/p.0.fpkm", "out/p.0.segments", "out/p.0.fraglength_effect"])
    import lzma                                   # 96 MB of SAM text: xz brings it to 3 MB (gzip: 6 MB)
    with open(f"{tmp}/in.sam", "rb") as f, open(os.path.join(HERE, "config1.in.sam.xz"), "wb") as g:
        g.write(lzma.compress(f.read(), preset=9 | lzma.PRESET_EXTREME))


def main():
    if len(sys.argv) > 1:          # python make_golden.py bigmod config1: only those
        for n in sys.argv[1:]:
            with tempfile.TemporaryDirectory() as tmp:
                globals()["fixture_" + n](tmp)
                print("ok", n)
        return
    if not (os.path.exists(REF) and os.path.exists(REFB)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    for fx in (fixture_se, fixture_pe, fixture_crafted, fixture_bowtie_pe, fixture_built, fixture_eumacut, fixture_bigmod, fixture_config1):
        with tempfile.TemporaryDirectory() as tmp:
            fx(tmp)
            print("ok", fx.__name__)
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE))
    print("golden dir: %.1f KB" % (tot / 1e3))


if __name__ == "__main__":
    main()
