"""Seeded synthetic rsh indices and read sets (SURVEY.md §8(d) "Concrete synthetic inputs").

One generator serves the parity tests (small, also written as `.rsh` + bowtie/SAM text so the real
reference binary can be run on them) and bench.py (human-scale, packed arrays only).

Index conventions follow the reference's `.rsh` semantics (reference src/emsar_functions.c:2071-2130
writer, :1351-1510 reader): class ids are positional; cid 0..T-1 are the singleton classes in tid
order (cid == tid); multi-tid classes follow ordered by (cardinality, first tid, lexicographic rest);
a class is a sorted multiset of tids (duplicates allowed, :1792).
"""
from __future__ import annotations

import dataclasses
import os
import struct
import zlib
from typing import Optional

import numpy as np


@dataclasses.dataclass
class SynthIndex:
    T: int                      # transcripts (max_tid + 1)
    names: list                 # transcript names (len T); may be lazily generated
    class_ptr: np.ndarray       # int64[C+1]  CSR offsets over ALL classes (singletons first)
    class_tid: np.ndarray       # int32[nnz]
    euma: np.ndarray            # int32[C, nF]
    has_node: np.ndarray        # uint8[C]    0 = singleton line without EUMA ("\t\t\t": no node)
    min_fraglength: int         # header field 3 (overwrites -f, :1419)
    max_fraglength: int         # header field 4 (overwrites -F, :1420)
    readlength: int             # header field 5 (-1 for SE)
    max_t_size: int             # header field 2 (rshbucket_max_t_size)

    @property
    def C(self) -> int:
        return len(self.class_ptr) - 1

    @property
    def frag_min(self) -> int:   # Fraglengths.min (determine_fraglength_range :2471-2475)
        return max(self.min_fraglength, self.readlength)

    @property
    def frag_max(self) -> int:
        return max(self.max_fraglength, self.frag_min)

    @property
    def nF(self) -> int:
        return self.frag_max - self.frag_min + 1


@dataclasses.dataclass
class SynthReads:
    read_ptr: np.ndarray        # int64[n+1]
    read_tid: np.ndarray        # int32[sum k]  (unsorted within a read)
    read_fraglen: np.ndarray    # int32[n]
    true_class: Optional[np.ndarray] = None  # int64[n]  (-1 = matches no class); generator's own truth


def _powerlaw_k(rng, n, alpha, kmin, kmax):
    ks = np.arange(kmin, kmax + 1)
    p = ks.astype(np.float64) ** (-alpha)
    p /= p.sum()
    return rng.choice(ks, size=n, p=p)


def make_index(T=2000, n_multi=10000, alpha=2.4, kmax=99, nF=1, seed=0, module_cap=500,
               p_dup_tid=0.01, p_no_node=0.0, frag_min=None, readlength=-1,
               hubs=0, hub_classes=0) -> SynthIndex:
    """Gene-family structured index. Modules (connected components) stay <= module_cap transcripts
    because class members are always drawn inside one family block of consecutive tids."""
    rng = np.random.default_rng(seed)
    # families: consecutive blocks of transcripts, heavy-tailed sizes capped by module_cap
    fam_sizes = []
    left = T
    while left > 0:
        s = int(min(left, module_cap, max(1, round(rng.pareto(1.2) * 8 + 1 + rng.geometric(0.2)))))
        fam_sizes.append(s)
        left -= s
    fam_sizes = np.array(fam_sizes, dtype=np.int64)
    fam_start = np.concatenate([[0], np.cumsum(fam_sizes)[:-1]])
    # classes by cardinality
    ks = _powerlaw_k(rng, n_multi, alpha, 2, kmax)
    blocks = {}
    order_sizes = np.argsort(fam_sizes)
    sizes_sorted = fam_sizes[order_sizes]
    for k in np.unique(ks):
        n_want = int((ks == k).sum())
        n_k = n_want + n_want // 3 + 8      # oversample: duplicate keys are dropped below, then trimmed back
        # a class with repeated tid may need only k-1 distinct members; keep it simple: need k distinct
        lo = np.searchsorted(sizes_sorted, k, side="left")
        if lo >= len(sizes_sorted):
            continue  # no family large enough for this cardinality
        elig = order_sizes[lo:]
        w_f = fam_sizes[elig].astype(np.float64)
        fam = elig[rng.choice(len(elig), size=n_k, p=w_f / w_f.sum())]
        fsz = fam_sizes[fam]
        # window inside the family: mostly gene-sized neighbourhoods, sometimes a wider paralog neighbourhood
        wide = rng.random(n_k) < 0.1
        w = np.where(wide, k + 16 + rng.geometric(0.05, size=n_k), k + rng.geometric(0.25, size=n_k))
        w = np.maximum(np.minimum(w, fsz), k)
        start = fam_start[fam] + (rng.random(n_k) * (fsz - w + 1)).astype(np.int64)
        rows = np.empty((n_k, k), dtype=np.int64)
        step = max(1, (1 << 24) // max(int(w.max()), 1))        # bound the scratch to ~16M keys per chunk
        for c0 in range(0, n_k, step):
            wc = w[c0:c0 + step]
            wmax = int(wc.max())
            keys = rng.random((len(wc), wmax))
            keys[np.arange(wmax)[None, :] >= wc[:, None]] = 2.0  # outside the window: never chosen
            offs = np.argpartition(keys, k - 1, axis=1)[:, :k] if k < wmax else np.tile(np.arange(wmax), (len(wc), 1))[:, :k]
            rows[c0:c0 + step] = start[c0:c0 + step, None] + offs
        if p_dup_tid > 0:
            dup = rng.random(n_k) < p_dup_tid
            if dup.any():
                src = rng.integers(0, k, size=n_k)
                dst = (src + 1 + rng.integers(0, k - 1, size=n_k)) % k if k > 1 else src
                idx = np.nonzero(dup)[0]
                rows[idx, dst[idx]] = rows[idx, src[idx]]
        rows.sort(axis=1)
        rows = np.unique(rows, axis=0)          # drop duplicate keys; also gives canonical order
        if len(rows) > n_want:
            rows = rows[np.sort(rng.choice(len(rows), size=n_want, replace=False))]
        blocks[int(k)] = rows.astype(np.int32)
    # hub transcripts: `hubs` tids that each sit in `hub_classes` extra pair classes inside their family
    if hubs > 0 and hub_classes > 0:
        big = np.nonzero(fam_sizes >= min(200, int(fam_sizes.max())))[0]   # hubs need room for many distinct classes
        hub_f = big[rng.choice(len(big), size=hubs)]
        extra = []
        for f in hub_f:
            h = fam_start[f] + rng.integers(0, fam_sizes[f])
            other = fam_start[f] + rng.integers(0, fam_sizes[f], size=hub_classes)
            third = fam_start[f] + rng.integers(0, fam_sizes[f], size=hub_classes)
            r = np.stack([np.full(hub_classes, h), other, third], axis=1)
            extra.append(r)
        r = np.concatenate(extra)
        r.sort(axis=1)
        r = np.concatenate([blocks.get(3, np.zeros((0, 3), np.int32)), r.astype(np.int32)])
        blocks[3] = np.unique(r, axis=0)
    k_sorted = sorted(blocks)
    card = np.concatenate([np.full(len(blocks[k]), k, dtype=np.int64) for k in k_sorted]) if k_sorted else np.zeros(0, np.int64)
    multi_tid = np.concatenate([blocks[k].ravel() for k in k_sorted]) if k_sorted else np.zeros(0, np.int32)
    n_multi = len(card)
    C = T + n_multi
    class_ptr = np.zeros(C + 1, dtype=np.int64)
    class_ptr[1:T + 1] = np.arange(1, T + 1)
    class_ptr[T + 1:] = T + np.cumsum(card)
    class_tid = np.concatenate([np.arange(T, dtype=np.int32), multi_tid.astype(np.int32)])
    # EUMA (number of distinct k-mers / fragments per class and fragment length)
    a_single = np.floor(rng.lognormal(6.0, 0.8, size=T)).astype(np.int64) + 1
    # shared stretches get shorter as more transcripts share them
    a_multi = rng.geometric(np.minimum(0.9, 0.01 * card), size=n_multi).astype(np.int64)
    a = np.concatenate([a_single, a_multi])
    if nF == 1:
        euma = a[:, None].astype(np.int32)
    else:
        slope = a / (nF * rng.uniform(0.8, 3.0, size=C))
        f = np.arange(nF)[None, :]
        euma = np.maximum(0, np.floor(a[:, None] - slope[:, None] * f)).astype(np.int32)
    has_node = np.ones(C, dtype=np.uint8)
    if p_no_node > 0:
        nn = rng.random(T) < p_no_node
        has_node[:T][nn] = 0
        euma[:T][nn] = 0
    if frag_min is None:
        frag_min = 1 if readlength < 0 else readlength
    minf = frag_min
    maxf = frag_min + nF - 1
    max_t_size = int(card.max()) if n_multi else 1
    names = [f"T{t:07d}" for t in range(T)] if T <= 200000 else None
    return SynthIndex(T=T, names=names, class_ptr=class_ptr, class_tid=class_tid, euma=np.ascontiguousarray(euma),
                      has_node=has_node, min_fraglength=minf, max_fraglength=maxf, readlength=readlength,
                      max_t_size=max(max_t_size, 2))


# ----------------------------------------------------------------------------------------------
# Human-scale generators (bench.py, full-size tests): the loops run in emsar_b200/tools/synth_gen.c.
# ----------------------------------------------------------------------------------------------
_GEN = None


def _gen():
    global _GEN
    if _GEN is None:
        import ctypes as C
        so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libemsar_synth.so")
        if not os.path.exists(so):
            from . import build
            build.build_tools()
        _GEN = C.CDLL(so)
        _GEN.synth_reads_pass1.restype = C.c_int64
    return _GEN


def _cp(a):
    import ctypes as C
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_index_v2(T=200000, n_multi=1800000, alpha=2.4, kmax=99, nF=1, seed=0, module_cap=5000, p_cross=0.1, scatter=True,
                  p_dup_tid=0.01, hubs=0, hub_classes=0, frag_min=None, readlength=-1, shuffle_tids=False) -> SynthIndex:
    """SURVEY.md section 8(d) generator at full size. Gene families are blocks of consecutive tids; families are grouped into
    paralog families of at most `module_cap` transcripts (from scattered places of the tid range when `scatter`, so that the
    fasta order does NOT keep a module together); a class draws its members inside one gene family (1 - p_cross) or anywhere
    in its paralog family (p_cross). Connected modules therefore stay <= module_cap transcripts (EUMAcut = 0). `hubs` transcripts
    (spread over the largest paralog families) are forced into about `hub_classes` classes each. `shuffle_tids` finally renames
    the transcripts by a random permutation (a fasta in arbitrary order)."""
    import ctypes as C
    rng = np.random.default_rng(seed)
    G = _gen()
    fam_sizes = []
    left = T
    while left > 0:
        s = int(min(left, module_cap, max(1, round(rng.pareto(1.2) * 8 + 1 + rng.geometric(0.2)))))
        fam_sizes.append(s)
        left -= s
    fam_sizes = np.array(fam_sizes, dtype=np.int64)
    nfam = len(fam_sizes)
    fam_start = np.concatenate([[0], np.cumsum(fam_sizes)[:-1]])
    # paralog families: greedy bins of families (random order when scattered) with at most module_cap transcripts
    order = rng.permutation(nfam) if scatter else np.arange(nfam)
    grp_of_fam = np.zeros(nfam, dtype=np.int64)
    g, acc = 0, 0
    for f in order:
        if acc + fam_sizes[f] > module_cap and acc > 0:
            g += 1
            acc = 0
        grp_of_fam[f] = g
        acc += fam_sizes[f]
    ngrp = g + 1
    # virtual order: families of one paralog family are contiguous; vmap[virtual position] = tid
    forder = np.lexsort((np.arange(nfam), grp_of_fam))
    vfam_start = np.zeros(nfam, dtype=np.int64)
    vfam_start[forder] = np.concatenate([[0], np.cumsum(fam_sizes[forder])[:-1]])
    vmap = np.empty(T, dtype=np.int32)
    for f in range(nfam):
        vmap[vfam_start[f]:vfam_start[f] + fam_sizes[f]] = np.arange(fam_start[f], fam_start[f] + fam_sizes[f], dtype=np.int32)
    grp_size = np.bincount(grp_of_fam, weights=fam_sizes, minlength=ngrp).astype(np.int64)
    grp_vstart = np.zeros(ngrp, dtype=np.int64)
    np.minimum.at(grp_vstart, grp_of_fam, np.iinfo(np.int64).max)
    grp_vstart[:] = np.iinfo(np.int64).max
    np.minimum.at(grp_vstart, grp_of_fam, vfam_start)
    # hubs: spread over the largest paralog families
    hub_ptr = hub_tid = hub_p = None
    ks = _powerlaw_k(rng, n_multi, alpha, 2, kmax)
    if hubs > 0 and hub_classes > 0:
        big = np.argsort(-grp_size)[:max(1, min(ngrp, (hubs + 3) // 4))]
        hg = big[np.arange(hubs) % len(big)]
        ht = np.array([vmap[grp_vstart[g_] + rng.integers(0, grp_size[g_])] for g_ in hg], dtype=np.int32)
        o = np.argsort(hg, kind="stable")
        hg, ht = hg[o], ht[o]
        hub_ptr = np.zeros(ngrp + 1, dtype=np.int64)
        np.add.at(hub_ptr, hg + 1, 1)
        hub_ptr = np.cumsum(hub_ptr)
        hub_tid = np.ascontiguousarray(ht)
        exp_cls = n_multi * grp_size / float(T)                 # classes a paralog family receives (drawn in proportion to its size)
        hub_p = np.minimum(0.9, 1.4 * hub_classes / np.maximum(exp_cls, 1.0))
    blocks = {}
    order_sizes = np.argsort(fam_sizes)
    sizes_sorted = fam_sizes[order_sizes]
    gorder = np.argsort(grp_size)
    gsorted = grp_size[gorder]
    for k in np.unique(ks):
        k = int(k)
        n_want = int((ks == k).sum())
        n_k = int(n_want * (2.2 if k == 2 else 1.5 if k == 3 else 1.34)) + 8      # duplicate keys are dropped below (small genes collide often)
        lo_f = np.searchsorted(sizes_sorted, k, side="left")
        lo_g = np.searchsorted(gsorted, k, side="left")
        if lo_g >= ngrp:
            continue
        cross = rng.random(n_k) < p_cross
        if lo_f >= nfam:
            cross[:] = True                      # no single gene family is large enough
        vlo = np.zeros(n_k, dtype=np.int64)
        w = np.zeros(n_k, dtype=np.int64)
        grp = np.zeros(n_k, dtype=np.int32)
        nin = int((~cross).sum())
        if nin:
            elig = order_sizes[lo_f:]
            wf = fam_sizes[elig].astype(np.float64)
            fam = elig[rng.choice(len(elig), size=nin, p=wf / wf.sum())]
            fsz = fam_sizes[fam]
            wide = rng.random(nin) < 0.1
            ww = np.where(wide, k + 16 + rng.geometric(0.05, size=nin), k + rng.geometric(0.25, size=nin))
            ww = np.maximum(np.minimum(ww, fsz), k)
            vlo[~cross] = vfam_start[fam] + (rng.random(nin) * (fsz - ww + 1)).astype(np.int64)
            w[~cross] = ww
            grp[~cross] = grp_of_fam[fam]
        nx = n_k - nin
        if nx:
            elig = gorder[lo_g:]
            wg = grp_size[elig].astype(np.float64)
            gg = elig[rng.choice(len(elig), size=nx, p=wg / wg.sum())]
            vlo[cross] = grp_vstart[gg]
            w[cross] = grp_size[gg]
            grp[cross] = gg
        rows = np.empty((n_k, k), dtype=np.int32)
        bad = G.synth_gen_classes(C.c_int64(n_k), C.c_int32(k), _cp(vlo), _cp(w), _cp(vmap), C.c_uint64(seed * 1000003 + k),
                                  C.c_double(p_dup_tid), _cp(grp), _cp(hub_ptr), _cp(hub_tid), _cp(hub_p), _cp(rows))
        assert bad == 0
        rows = np.unique(rows, axis=0)
        if len(rows) > n_want:
            rows = rows[np.sort(rng.choice(len(rows), size=n_want, replace=False))]
        blocks[k] = rows
    if shuffle_tids:
        perm = rng.permutation(T).astype(np.int32)        # new name of every transcript
        for k in list(blocks):
            r = perm[blocks[k]]
            r.sort(axis=1)
            blocks[k] = np.unique(r, axis=0)              # canonical class order under the new names
    k_sorted = sorted(blocks)
    card = np.concatenate([np.full(len(blocks[k]), k, dtype=np.int64) for k in k_sorted])
    multi_tid = np.concatenate([blocks[k].ravel() for k in k_sorted])
    n_multi = len(card)
    Cn = T + n_multi
    class_ptr = np.zeros(Cn + 1, dtype=np.int64)
    class_ptr[1:T + 1] = np.arange(1, T + 1)
    class_ptr[T + 1:] = T + np.cumsum(card)
    class_tid = np.concatenate([np.arange(T, dtype=np.int32), multi_tid.astype(np.int32)])
    a_single = np.floor(rng.lognormal(6.0, 0.8, size=T)).astype(np.int64) + 1
    a_multi = rng.geometric(np.minimum(0.9, 0.01 * card), size=n_multi).astype(np.int64)
    a = np.concatenate([a_single, a_multi]).astype(np.float64)
    if nF == 1:
        euma = a[:, None].astype(np.int32)
    else:
        slope = a / (nF * rng.uniform(0.8, 3.0, size=Cn))
        euma = np.empty((Cn, nF), dtype=np.int32)
        G.synth_euma_fill(C.c_int64(Cn), C.c_int32(nF), _cp(a), _cp(slope), _cp(euma))
    if frag_min is None:
        frag_min = 1 if readlength < 0 else readlength
    return SynthIndex(T=T, names=None, class_ptr=class_ptr, class_tid=class_tid, euma=euma, has_node=np.ones(Cn, dtype=np.uint8),
                      min_fraglength=frag_min, max_fraglength=frag_min + nF - 1, readlength=readlength,
                      max_t_size=max(int(card.max()) if n_multi else 1, 2))


def make_reads_fast(idx: SynthIndex, N, seed=0, p_zero=0.3, p_unmatched=0.005) -> SynthReads:
    """make_reads for 10^7..10^8 reads: the class of every read is an independent draw from p_c (alias method), which has the
    distribution of the multinomial-then-shuffle of make_reads; tid lists rotated / reversed, unmatched pairs and fragment
    lengths as there. Threads: OpenMP; the result depends on (idx, N, seed) only."""
    import ctypes as C
    G = _gen()
    p = class_rates(idx, true_theta(idx, seed, p_zero))
    Cn = idx.C
    prob = np.empty(Cn, dtype=np.float64)
    alias = np.empty(Cn, dtype=np.int64)
    rc = G.synth_alias_build(C.c_int64(Cn), _cp(np.ascontiguousarray(p)), _cp(prob), _cp(alias))
    assert rc == 0
    N = int(N)
    cls = np.empty(N, dtype=np.int64)
    read_ptr = np.empty(N + 1, dtype=np.int64)
    cp = np.ascontiguousarray(idx.class_ptr, dtype=np.int64)
    ct = np.ascontiguousarray(idx.class_tid, dtype=np.int32)
    tot = G.synth_reads_pass1(C.c_int64(N), C.c_uint64(seed), C.c_int64(Cn), _cp(cp), _cp(prob), _cp(alias), C.c_double(p_unmatched),
                              _cp(cls), _cp(read_ptr))
    read_tid = np.empty(int(tot), dtype=np.int32)
    fl = np.empty(N, dtype=np.int32)
    cdf = np.ascontiguousarray(np.cumsum(frag_weights(idx)))
    G.synth_reads_pass2(C.c_int64(N), C.c_uint64(seed), C.c_int32(idx.T), _cp(cp), _cp(ct), _cp(cls), _cp(read_ptr), C.c_int32(idx.nF),
                        C.c_int32(idx.frag_min), _cp(cdf), _cp(read_tid), _cp(fl))
    return SynthReads(read_ptr=read_ptr, read_tid=read_tid, read_fraglen=fl, true_class=cls)


def true_theta(idx: SynthIndex, seed=0, p_zero=0.3):
    rng = np.random.default_rng(seed + 7919)
    th = rng.lognormal(0.0, 2.0, size=idx.T)
    th[rng.random(idx.T) < p_zero] = 0.0
    return th


def frag_weights(idx: SynthIndex):
    nF = idx.nF
    if nF == 1:
        return np.ones(1)
    f = np.arange(nF) + idx.frag_min
    mu = idx.frag_min + 0.5 * nF
    w = np.exp(-0.5 * ((f - mu) / (0.1 * nF + 1.0)) ** 2)
    return w / w.sum()


def class_rates(idx: SynthIndex, theta):
    """p_c ∝ (Σ_f w_f·EUMA[c][f]) · Σ_{t∈c} θ_t   (reference model: lambdap :2966-2975)."""
    w = frag_weights(idx)
    a = idx.euma.astype(np.float64) @ w
    a = a * idx.has_node
    seg = np.add.reduceat(theta[idx.class_tid], idx.class_ptr[:-1].astype(np.int64))
    return a * seg


def sample_class_counts(idx: SynthIndex, N, seed=0, p_zero=0.3):
    rng = np.random.default_rng(seed + 104729)
    p = class_rates(idx, true_theta(idx, seed, p_zero))
    p = p / p.sum()
    return rng.multinomial(N, p).astype(np.int64)


def make_reads(idx: SynthIndex, N, seed=0, p_zero=0.3, p_unmatched=0.005, shuffle_tids=True) -> SynthReads:
    """Read tid-lists as the host reader would hand them to update_ReadCounts (:838): one list per read
    group, unsorted, plus the group's fragment length. `p_unmatched` of the reads carry a tid multiset that
    is (almost surely) not a class: they count toward N and the fragment histogram only (:940-941)."""
    rng = np.random.default_rng(seed + 15485863)
    n_un = int(round(N * p_unmatched))
    counts = sample_class_counts(idx, N - n_un, seed, p_zero)
    cls = np.repeat(np.arange(idx.C, dtype=np.int64), counts)
    cls = np.concatenate([cls, np.full(n_un, -1, dtype=np.int64)])
    rng.shuffle(cls)
    n = len(cls)
    k = np.where(cls >= 0, idx.class_ptr[np.maximum(cls, 0) + 1] - idx.class_ptr[np.maximum(cls, 0)], 2)
    read_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(k, out=read_ptr[1:])
    tot = int(read_ptr[-1])
    # ragged gather
    rid = np.repeat(np.arange(n, dtype=np.int64), k)
    pos = np.arange(tot, dtype=np.int64) - read_ptr[rid]
    kk = k[rid]
    if shuffle_tids:
        rot = rng.integers(0, 1 << 30, size=n)[rid]
        rev = (rng.random(n) < 0.5)[rid]
        pos = (pos + rot) % kk
        pos = np.where(rev, kk - 1 - pos, pos)
    src = idx.class_ptr[np.maximum(cls, 0)][rid] + pos
    read_tid = idx.class_tid[src].astype(np.int32)
    un = np.nonzero(cls < 0)[0]
    if len(un):
        # two far-apart random tids: virtually never a class (classes live inside one family block)
        a = rng.integers(0, idx.T, size=len(un))
        b = (a + idx.T // 2 + rng.integers(0, max(1, idx.T // 4), size=len(un))) % idx.T
        read_tid[read_ptr[un]] = a
        read_tid[read_ptr[un] + 1] = b
    if idx.nF == 1:
        fl = np.full(n, idx.frag_min, dtype=np.int32)
    else:
        fl = (idx.frag_min + rng.choice(idx.nF, size=n, p=frag_weights(idx))).astype(np.int32)
    return SynthReads(read_ptr=read_ptr, read_tid=read_tid, read_fraglen=fl, true_class=cls)


# ----------------------------------------------------------------------------------------------
# Text writers (small fixtures only): `.rsh` and bowtie / SAM alignments for the reference binary.
# ----------------------------------------------------------------------------------------------

def write_rsh(idx: SynthIndex, path: str):
    """`.rsh` text exactly as print_rsh (:2071-2130) lays it out (trailing commas included)."""
    with open(path, "w") as f:
        f.write(f"#{idx.T - 1},{idx.max_t_size},{idx.min_fraglength},{idx.max_fraglength},{idx.readlength}\n")
        for t in range(idx.T):
            f.write(f"@{t}\t{idx.names[t]}\n")
        f.write("cid\tno.tids\tfirst.tid\tother.tids\tsegment.length\n")
        cp, ct = idx.class_ptr, idx.class_tid
        for c in range(idx.C):
            tids = ct[cp[c]:cp[c + 1]]
            k = len(tids)
            if k == 1 and not idx.has_node[c]:
                f.write(f"{c}\t1\t{tids[0]}\t\t\t\n")
                continue
            e = "".join(f"{v}," for v in idx.euma[c])
            if k == 1:
                f.write(f"{c}\t1\t{tids[0]}\t\t{e}\n")
            else:
                o = "".join(f"{v}," for v in tids[1:])
                f.write(f"{c}\t{k}\t{tids[0]}\t{o}\t{e}\n")


def write_bowtie_se(idx: SynthIndex, reads: SynthReads, path: str):
    """Default bowtie output, SE: id, strand, tname, pos, seq, qual, other, mismatches (parse_bowtieline
    :552-587). The fragment length is strlen(seq) (:572), so `A`*fraglen keeps lines short."""
    rp, rt, fl = reads.read_ptr, reads.read_tid, reads.read_fraglen
    with open(path, "w") as f:
        for r in range(len(fl)):
            seq = "A" * int(fl[r])
            for j, t in enumerate(rt[rp[r]:rp[r + 1]]):
                # distinct pos per alignment so that the duplicate filter (alignment.c:36-40) keeps repeats
                f.write(f"r{r}\t+\t{idx.names[t]}\t{j}\t{seq}\t{seq}\t0\t\n")


def write_sam_pe(idx: SynthIndex, reads: SynthReads, path: str, tlen=100000):
    """SAM, PE: mates adjacent, flags 0x40/0x80 (+0x10 on the reverse mate), MD:Z tags, l_qseq = readlength,
    fraglen = pos2 - pos1 + readlength (convert_bam_alignment_2_alignment_PE :426-469)."""
    L = idx.readlength
    rp, rt, fl = reads.read_ptr, reads.read_tid, reads.read_fraglen
    seq = "A" * L
    with open(path, "w") as f:
        f.write("@HD\tVN:1.0\tSO:unsorted\n")
        for t in range(idx.T):
            f.write(f"@SQ\tSN:{idx.names[t]}\tLN:{tlen}\n")
        for r in range(len(fl)):
            d = int(fl[r]) - L
            for j, t in enumerate(rt[rp[r]:rp[r + 1]]):
                p1 = 1 + 1000 * j      # SAM is 1-based; distinct pos per alignment
                p2 = p1 + d
                n = idx.names[t]
                if d > 0:
                    f.write(f"r{r}\t{0x1 | 0x2 | 0x20 | 0x40}\t{n}\t{p1}\t255\t{L}M\t=\t{p2}\t{d + L}\t{seq}\t{seq}\tMD:Z:{L}\n")
                    f.write(f"r{r}\t{0x1 | 0x2 | 0x10 | 0x80}\t{n}\t{p2}\t255\t{L}M\t=\t{p1}\t{-(d + L)}\t{seq}\t{seq}\tMD:Z:{L}\n")
                else:
                    # pos2 == pos1 is treated as "mate2(f)...mate1(r)" (:461-465)
                    f.write(f"r{r}\t{0x1 | 0x2 | 0x10 | 0x40}\t{n}\t{p1}\t255\t{L}M\t=\t{p2}\t{L}\t{seq}\t{seq}\tMD:Z:{L}\n")
                    f.write(f"r{r}\t{0x1 | 0x2 | 0x20 | 0x80}\t{n}\t{p2}\t255\t{L}M\t=\t{p1}\t{-L}\t{seq}\t{seq}\tMD:Z:{L}\n")


def sam_to_bam(sam_path, bam_path):
    """Minimal BAM writer (BGZF blocks over zlib) for the fixtures; records carry what the reference reads."""
    refs, recs, text = [], [], []
    for line in open(sam_path):
        line = line.rstrip("\n")
        if line.startswith("@"):
            text.append(line)
            if line.startswith("@SQ"):
                d = dict(x.split(":", 1) for x in line.split("\t")[1:])
                refs.append((d["SN"], int(d["LN"])))
            continue
        recs.append(line.split("\t"))
    ref_id = {n: i for i, (n, _) in enumerate(refs)}
    out = bytearray()
    htext = ("\n".join(text) + "\n").encode()
    out += b"BAM\1" + struct.pack("<i", len(htext)) + htext + struct.pack("<i", len(refs))
    for n, ln in refs:
        out += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", ln)
    codes = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    for f in recs:
        qname, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = f[:11]
        tags = f[11:]
        rid = ref_id.get(rname, -1)
        nid = rid if rnext == "=" else ref_id.get(rnext, -1)
        cig = []
        num = ""
        for ch in cigar:
            if ch.isdigit():
                num += ch
            elif ch != "*":
                cig.append((int(num) << 4) | "MIDNSHP=X".index(ch))
                num = ""
        l_seq = 0 if seq == "*" else len(seq)
        sq = bytearray((l_seq + 1) // 2)
        for i, ch in enumerate(seq if seq != "*" else ""):
            sq[i // 2] |= codes.get(ch, 15) << (4 if i % 2 == 0 else 0)
        ql = bytes([0xFF] * l_seq) if qual == "*" else bytes(ord(c) - 33 for c in qual)
        aux = bytearray()
        for t in tags:
            tag, ty, val = t.split(":", 2)
            if ty == "Z":
                aux += tag.encode() + b"Z" + val.encode() + b"\0"
            elif ty == "i":
                aux += tag.encode() + b"i" + struct.pack("<i", int(val))
        body = struct.pack("<iiBBHHHiiii", rid, int(pos) - 1, len(qname) + 1, int(mapq), 4680, len(cig), int(flag), l_seq, nid,
                           int(pnext) - 1, int(tlen))
        body += qname.encode() + b"\0" + b"".join(struct.pack("<I", c) for c in cig) + bytes(sq) + ql + bytes(aux)
        out += struct.pack("<i", len(body)) + body
    with open(bam_path, "wb") as g:
        for o in list(range(0, len(out), 60000)) + [None]:
            chunk = b"" if o is None else bytes(out[o:o + 60000])
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            cd = co.compress(chunk) + co.flush()
            bsize = len(cd) + 25
            g.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + cd +
                    struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))


def ensure_dir(p):
    os.makedirs(p, exist_ok=True)
    return p
