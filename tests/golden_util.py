"""Helpers for the golden fixtures under tests/golden/ (made by tests/golden/make_golden.py from the unmodified
reference binaries)."""
import gzip
import os
import shutil

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (alignment file, format, pe, strand, -k)
FIXTURES = {
    "se": dict(aln="se.in.bowtie", rsh="se.in.rsh", fmt="bowtie", pe=False, strand="ns", k=100, out="se", rounds=8),
    "pe": dict(aln="pe.in.sam", rsh="pe.in.rsh", fmt="sam", pe=True, strand="ns", k=100, out="pe", rounds=8),
    "pe_bam_ssfr": dict(aln="pe.in.bam", rsh="pe.in.rsh", fmt="bam", pe=True, strand="ssfr", k=100, out="pe_bam_ssfr", rounds=2),
    "crafted": dict(aln="crafted.in.bowtie", rsh="crafted.in.rsh", fmt="bowtie", pe=False, strand="ssf", k=3, out="crafted", rounds=2),
    "built": dict(aln="built.in.sam", rsh="built.in.rsh", fmt="sam", pe=False, strand="ssf", k=100, out="built", rounds=8),
    # two sequence-sharing sets of ~700 transcripts: the reference's rounds visibly disagree here (the 6 sd term of the tolerance policy)
    "bigmod": dict(aln="bigmod.in.bowtie", rsh="bigmod.in.rsh", fmt="bowtie", pe=False, strand="ns", k=100, out="bigmod", rounds=8),
    # substitute of BASELINE configs[0] (the bundled Vicugna PE sample is absent): index by the reference's emsar-build -P on a generated
    # 2K-transcript fasta, 100K simulated PE fragments as SAM
    "config1": dict(aln="config1.in.sam", rsh="config1.in.rsh", fmt="sam", pe=True, strand="ns", k=100, out="config1", rounds=8),
    "bowtie_pe_ns": dict(aln="bowtie_pe.in.bowtie", rsh="bowtie_pe.in.rsh", fmt="bowtie", pe=True, strand="ns", k=100, out="bowtie_pe_ns", rounds=2),
    "bowtie_pe_ssfr": dict(aln="bowtie_pe.in.bowtie", rsh="bowtie_pe.in.rsh", fmt="bowtie", pe=True, strand="ssfr", k=100, out="bowtie_pe_ssfr", rounds=2),
    "bowtie_pe_ssrf": dict(aln="bowtie_pe.in.bowtie", rsh="bowtie_pe.in.rsh", fmt="bowtie", pe=True, strand="ssrf", k=100, out="bowtie_pe_ssrf", rounds=2),
}


def materialize(name, tmpdir):
    """Gunzip a fixture file into tmpdir and return its path (the .bam is stored as is)."""
    src = os.path.join(GOLD, name)
    dst = os.path.join(str(tmpdir), name)
    if os.path.exists(src):
        shutil.copy(src, dst)
    elif os.path.exists(src + ".xz"):
        import lzma
        with lzma.open(src + ".xz", "rb") as f, open(dst, "wb") as g:
            shutil.copyfileobj(f, g)
    else:
        with gzip.open(src + ".gz", "rb") as f, open(dst, "wb") as g:
            shutil.copyfileobj(f, g)
    return dst


def _rows(name):
    with gzip.open(os.path.join(GOLD, name + ".gz"), "rt") as f:
        return [l.rstrip("\n").split("\t") for l in f][1:]


def read_segments(prefix):
    rows = _rows(prefix + ".p.0.segments")
    return dict(set_id=np.array([int(r[1][1:]) for r in rows], dtype=np.int32), adjEUMA=np.array([float(r[4]) for r in rows]),
                ReadCount=np.array([int(r[5]) for r in rows], dtype=np.int32), expected=np.array([float(r[6]) for r in rows]),
                text=rows)


def read_fraglength(prefix):
    rows = _rows(prefix + ".p.0.fraglength_effect")
    return dict(length=np.array([int(r[0]) for r in rows]), counts=np.array([int(r[1]) for r in rows], dtype=np.int64),
                Wf=np.array([float(r[2]) for r in rows]), text=rows)


def has_fpkm(prefix):
    return os.path.exists(os.path.join(GOLD, prefix + ".p.0.fpkm.gz"))


def read_fpkm(prefix):
    rows = _rows(prefix + ".p.0.fpkm")
    f = lambda i: np.array([float(r[i]) for r in rows])
    return dict(names=[r[0] for r in rows], fpkm=f(1), sd=f(2), efflen=f(3), ireadcount=f(4),
                ireadcount_int=np.array([int(r[5]) for r in rows]), tpm=f(6))


def parse_out_file(path):
    with open(path) as f:
        return [l.rstrip("\n").split("\t") for l in f][1:]


PRINT_EPS = 5e-7          # the reference's files carry 6 decimals (%lf): a printed value is within 5e-7 of the number behind it


def fpkm_tolerance(gold_fpkm, efflen, N, rounds, from_files=False):
    """SURVEY.md §8(c) item 3 / north star: |x - m| <= max(1e-6 |m|, 1e-3 reads-equivalent, 6 s_t), nothing else - except that m (and s_t) are
    read from a file with 6 decimals, so the rounding of the printed numbers (5e-7 each; twice that when x is read from a file too) is added."""
    m = gold_fpkm["fpkm"]
    s_t = gold_fpkm["sd"] * rounds            # the file stores sd / NUM_ROUND (:3200)
    s_t = np.where(np.isfinite(s_t), s_t + PRINT_EPS * rounds, 0.0)
    per_read = np.where(efflen > 0, 1e-3 / np.maximum(efflen / 1e3 * N / 1e6, 1e-300), np.inf)
    return np.maximum(np.maximum(1e-6 * np.abs(m), 6 * s_t), per_read) + PRINT_EPS * (2 if from_files else 1)


def ireadcount_tolerance(tol_fpkm, efflen, N, from_files=False):
    """iReadcount_t = iEUMA_t / 1e3 * FPKM_t * N / 1e6 (print_FPKMfinal :3203): the FPKM tolerance carried through that product."""
    return np.where(efflen > 0, np.where(np.isfinite(tol_fpkm), tol_fpkm, 0.0) * (efflen / 1e3 * N / 1e6), 0.0) + PRINT_EPS * (2 if from_files else 1)


def expected_tolerance(tol_fpkm, class_ptr, class_tid, adjEUMA, N, from_files=False):
    """expected_Readcount_c = sum_{t in c} FPKM_t * adjEUMA_c / 1e3 * N / 1e6 (print_aEUMA_3 :2289-2296): the FPKM tolerances of the members
    carried through that sum. Transcripts without a finite tolerance (no effective length) contribute through the class's own 1e-3 reads."""
    t = np.where(np.isfinite(tol_fpkm), tol_fpkm, 0.0)
    seg = np.add.reduceat(t[class_tid], np.asarray(class_ptr[:-1], dtype=np.int64))
    return np.maximum(seg * (adjEUMA / 1e3 * N / 1e6), 1e-3) + PRINT_EPS * (2 if from_files else 1)
