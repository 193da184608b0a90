"""The device builder's interface (include/emsar_cuda.h: emsar_build_desc / emsar_build_classes) pinned on the CPU: emsar_rsh_build of
libemsar_host.so is handed a `device_run` callback written HERE in plain Python straight from the header's description (group equal
read-length windows / mate pairs, singles, classes with their counts) and must produce the index the host construction - and therefore the
reference's emsar-build - writes, byte for byte. This checks what the host passes (sequence layout, starts, distance range) and how it folds
the answer (build_index.c::build_device), without a GPU; tests/test_build_index_gpu.py checks the CUDA implementation of the same interface."""
import collections
import ctypes as C
import gzip
import os

import numpy as np
import pytest

from test_build_index_cpu import GOLD, make_fasta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Desc(C.Structure):
    _fields_ = [("seq", C.POINTER(C.c_char)), ("border", C.c_int64), ("end", C.c_int64), ("T", C.c_int32), ("start", C.POINTER(C.c_int64)),
                ("pe", C.c_int32), ("stranded", C.c_int32), ("readlen", C.c_int32), ("d_min", C.c_int32), ("d_max", C.c_int32), ("max_repeat", C.c_int32)]


class Classes(C.Structure):
    _fields_ = [("T", C.c_int32), ("n_d", C.c_int32), ("single_count", C.POINTER(C.c_int32)), ("n_class", C.c_int64), ("class_off", C.POINTER(C.c_int64)),
                ("class_tid", C.POINTER(C.c_int32)), ("class_d", C.POINTER(C.c_int32)), ("class_count", C.POINTER(C.c_int32)),
                ("occurrences", C.c_int64), ("runs", C.c_int64), ("partitions", C.c_int32), ("device_ms", C.c_double), ("owner", C.c_void_p)]


RUN_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Desc), C.POINTER(Classes))
FREE_FN = C.CFUNCTYPE(None, C.POINTER(Classes))
ERR_FN = C.CFUNCTYPE(C.c_void_p)          # const char *(*)(void): the address of a buffer kept alive below
_MSG1 = C.create_string_buffer(b"python stand-in failed")
_MSG2 = C.create_string_buffer(b"two different substrings share a 128-bit hash")


class Opts(C.Structure):
    _fields_ = [("pe", C.c_int), ("stranded", C.c_int), ("readlength", C.c_int), ("readlen_min", C.c_int), ("readlen_max", C.c_int),
                ("min_fraglength", C.c_int), ("max_fraglength", C.c_int), ("max_repeat", C.c_int), ("header", C.c_char), ("threads", C.c_int),
                ("device_run", RUN_FN), ("device_free", FREE_FN), ("device_error", ERR_FN), ("device_ctx", C.c_void_p)]


def brute_force(d, keep):
    """emsar_build_classes from emsar_build_desc, the slow and obvious way"""
    n = d.end + 1
    S = C.string_at(d.seq, n)
    L, T = d.readlen, d.T
    start = [d.start[t] for t in range(T + 1)]
    ok = [i + L <= n and all(c in b"ACGT" for c in S[i:i + L]) for i in range(n)]
    dmin, dmax = (d.d_min, d.d_max) if d.pe else (0, 0)
    nD = dmax - dmin + 1
    groups = collections.defaultdict(list)
    tid_of = np.searchsorted(np.array(start[1:]), np.arange(d.border), side="right")
    for i in range(d.border):
        if not ok[i]:
            continue
        t = int(tid_of[i])
        if not d.pe:
            w = S[i:i + L]
            if not d.stranded:
                fl = d.end - i - L
                if w > S[fl:fl + L]:
                    w = S[fl:fl + L]
            groups[w].append((t, 0))
            continue
        for p in ([i] if d.stranded else [i, d.end - i - L]):
            if p < d.border:
                lo, hi = start[t], start[t + 1] - 1 - L
            else:
                lo, hi = d.end - (start[t + 1] - 1), d.end - start[t] - L
            for dd in range(dmin, dmax + 1):
                q = p + dd
                if q < lo or q > hi or not ok[q]:
                    continue
                if not d.stranded:
                    e = d.end - q - L
                    a, b = (S[p:p + L], S[q:q + L]), (S[e:e + L], S[e + dd:e + dd + L])
                    if not ((p < d.border and a <= b) or (p > d.border and a < b)):
                        continue
                groups[(S[p:p + L], S[q:q + L])].append((t, dd))
    single = np.zeros(T * nD, dtype=np.int32)
    classes = collections.Counter()
    for members in groups.values():
        if len(members) == 1:
            single[members[0][0] * nD + members[0][1] - dmin] += 1
        elif len(members) < d.max_repeat and len({m[1] for m in members}) == 1:
            classes[(tuple(sorted(m[0] for m in members)), members[0][1] - dmin)] += 1
    keys = list(classes)
    off = np.zeros(len(keys) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(k[0]) for k in keys])
    tids = np.array([t for k in keys for t in k[0]] or [0], dtype=np.int32)
    cd = np.array([k[1] for k in keys] or [0], dtype=np.int32)
    cnt = np.array([classes[k] for k in keys] or [0], dtype=np.int32)
    keep.append((single, off, tids, cd, cnt))
    return T, nD, single, len(keys), off, tids, cd, cnt, sum(len(m) for m in groups.values()), len(groups)


CASES = {"se_ns": dict(readlen_min=25, readlen_max=25), "se_ssf": dict(readlen_min=25, readlen_max=25, stranded=1),
         "se_range_k5": dict(readlen_min=24, readlen_max=27, max_repeat=5),
         "pe_ns": dict(pe=1, readlength=25, min_fraglength=40, max_fraglength=70),
         "pe_ssrf": dict(pe=1, stranded=1, readlength=20, min_fraglength=30, max_fraglength=55),
         "pe_refseq_k4": dict(pe=1, readlength=22, min_fraglength=1, max_fraglength=45, max_repeat=4, header=b"R")}


@pytest.mark.parametrize("name", list(CASES))
def test_host_fold_of_the_device_interface(built, tmp_path, name):
    from emsar_b200 import host
    L = host.lib()
    fa = str(tmp_path / "t.fa")
    make_fasta(fa, refseq="refseq" in name)
    keep, calls = [], []

    def run(ctx, dp, outp):
        d = dp.contents
        calls.append((d.readlen, d.pe, d.stranded, d.d_min, d.d_max))
        T, nD, single, nc, off, tids, cd, cnt, occ, runs = brute_force(d, keep)
        o = outp.contents
        o.T, o.n_d, o.n_class, o.occurrences, o.runs, o.partitions = T, nD, nc, occ, runs, 1
        o.single_count = single.ctypes.data_as(C.POINTER(C.c_int32))
        o.class_off = off.ctypes.data_as(C.POINTER(C.c_int64))
        o.class_tid = tids.ctypes.data_as(C.POINTER(C.c_int32))
        o.class_d = cd.ctypes.data_as(C.POINTER(C.c_int32))
        o.class_count = cnt.ctypes.data_as(C.POINTER(C.c_int32))
        return 0

    o = Opts()
    kw = dict(min_fraglength=1, max_fraglength=400, max_repeat=100, header=b"E")
    kw.update(CASES[name])
    for k, v in kw.items():
        setattr(o, k, v)
    o.device_run, o.device_free, o.device_error = RUN_FN(run), FREE_FN(lambda p: None), ERR_FN(lambda: C.addressof(_MSG1))
    rsh = C.POINTER(host._Rsh)()
    err = C.create_string_buffer(host.ERRLEN)
    L.emsar_rsh_build.argtypes = [C.c_char_p, C.POINTER(Opts), C.POINTER(C.POINTER(host._Rsh)), C.c_char_p]
    assert L.emsar_rsh_build(os.fsencode(fa), C.byref(o), C.byref(rsh), err) == 0, err.value
    out = str(tmp_path / "x.rsh")
    assert L.emsar_rsh_write(rsh, int(kw.get("pe", 0)), os.fsencode(out), err) == 0, err.value
    L.emsar_rsh_free(rsh)
    assert open(out, "rb").read() == gzip.open(os.path.join(GOLD, f"build_{name}.rsh.gz"), "rb").read()
    n_pass = 1 if kw.get("pe") else kw["readlen_max"] - kw["readlen_min"] + 1
    assert len(calls) == n_pass
    if kw.get("pe"):
        fmin = max(kw["min_fraglength"], kw["readlength"])
        assert calls[0][3:] == (fmin - kw["readlength"], kw["max_fraglength"] - kw["readlength"])


def test_device_failure_is_an_error(built, tmp_path):
    from emsar_b200 import host
    L = host.lib()
    fa = str(tmp_path / "t.fa")
    make_fasta(fa)
    o = Opts()
    for k, v in dict(readlen_min=25, readlen_max=25, min_fraglength=1, max_fraglength=400, max_repeat=100, header=b"E").items():
        setattr(o, k, v)
    o.device_run, o.device_free, o.device_error = RUN_FN(lambda c, d, r: 7), FREE_FN(lambda p: None), ERR_FN(lambda: C.addressof(_MSG2))
    rsh = C.POINTER(host._Rsh)()
    err = C.create_string_buffer(host.ERRLEN)
    L.emsar_rsh_build.argtypes = [C.c_char_p, C.POINTER(Opts), C.POINTER(C.POINTER(host._Rsh)), C.c_char_p]
    assert L.emsar_rsh_build(os.fsencode(fa), C.byref(o), C.byref(rsh), err) != 0
    assert b"128-bit hash" in err.value
