"""ctypes front-end of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module (and only as the checker). The product path (emsar_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "emsar_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_loglik.restype = C.c_double
        _LIB.orc_em_time.restype = C.c_double
    return _LIB


def ref_binary(name="emsar"):
    """Path of the UNMODIFIED reference binary built by oracle/Makefile, or None."""
    p = os.path.join(_HERE, "_ref", name)
    if not os.path.exists(p) and os.path.isdir("/root/reference/src"):
        subprocess.call(["make", "-s", "-C", _HERE, "ref"])
    return p if os.path.exists(p) else None


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def filter_group(tid, mm, fraglen, pos, max_repeat=100, pe=False):
    tid, mm, fraglen, pos = map(_i32, (tid, mm, fraglen, pos))
    keep = np.zeros(max(1, len(tid)), dtype=np.int32)
    n = lib().orc_filter_group(len(tid), _p(tid, C.c_int), _p(mm, C.c_int), _p(fraglen, C.c_int), _p(pos, C.c_int),
                               int(max_repeat), int(bool(pe)), _p(keep, C.c_int))
    return None if n < 0 else keep[:n].copy()


def count(idx, reads):
    """-> (ReadCount int32[C], FraglengthCounts int32[max_fraglength+1], N)."""
    cp, ct = _i64(idx.class_ptr), _i32(idx.class_tid)
    hn = np.ascontiguousarray(idx.has_node, dtype=np.uint8)
    rp, rt, fl = _i64(reads.read_ptr), _i32(reads.read_tid), _i32(reads.read_fraglen)
    R = np.zeros(idx.C, dtype=np.int32)
    F = np.zeros(idx.max_fraglength + 1, dtype=np.int32)
    N = C.c_int64(0)
    rc = lib().orc_count(int(idx.T), C.c_int64(idx.C), _p(cp, C.c_int64), _p(ct, C.c_int32), _p(hn, C.c_uint8),
                         int(idx.max_t_size), int(idx.min_fraglength), int(idx.max_fraglength),
                         C.c_int64(len(fl)), _p(rp, C.c_int64), _p(rt, C.c_int32), _p(fl, C.c_int32),
                         _p(R, C.c_int32), _p(F, C.c_int32), C.byref(N))
    assert rc == 0
    return R, F, int(N.value)


def prepare(idx, F, N, delta=0.0):
    cp, ct = _i64(idx.class_ptr), _i32(idx.class_tid)
    hn = np.ascontiguousarray(idx.has_node, dtype=np.uint8)
    eu = _i32(idx.euma)
    Wf = np.zeros(idx.nF)
    adj = np.zeros(idx.C)
    ps = np.zeros(idx.C)
    iE = np.zeros(idx.T)
    F = _i32(F)
    lib().orc_prepare(int(idx.T), C.c_int64(idx.C), _p(cp, C.c_int64), _p(ct, C.c_int32), int(idx.nF), _p(eu, C.c_int32),
                      _p(hn, C.c_uint8), int(idx.frag_min), _p(F, C.c_int32), C.c_int64(N), C.c_double(delta),
                      _p(Wf, C.c_double), _p(adj, C.c_double), _p(ps, C.c_double), _p(iE, C.c_double))
    return Wf, adj, ps, iE


def components(idx, adjEUMA, eumacut=0.0, max_ntid=5000):
    cp, ct = _i64(idx.class_ptr), _i32(idx.class_tid)
    CS = np.zeros(idx.C, dtype=np.int32)
    TS = np.zeros(idx.T, dtype=np.int32)
    cut = C.c_double(eumacut)
    adj = np.ascontiguousarray(adjEUMA, dtype=np.float64)
    ms = lib().orc_components(int(idx.T), C.c_int64(idx.C), _p(cp, C.c_int64), _p(ct, C.c_int32), _p(adj, C.c_double),
                              C.byref(cut), int(max_ntid), _p(CS, C.c_int32), _p(TS, C.c_int32))
    return ms, float(cut.value), CS, TS


def loglik(idx, R, EUMAps, theta, in_model=None):
    cp, ct = _i64(idx.class_ptr), _i32(idx.class_tid)
    R = _i32(R)
    ps = np.ascontiguousarray(EUMAps, dtype=np.float64)
    th = np.ascontiguousarray(theta, dtype=np.float64)
    im = None if in_model is None else np.ascontiguousarray(in_model, dtype=np.uint8)
    return float(lib().orc_loglik(C.c_int64(idx.C), _p(cp, C.c_int64), _p(ct, C.c_int32), _p(R, C.c_int32),
                                  _p(ps, C.c_double), _p(im, C.c_uint8) if im is not None else None, _p(th, C.c_double)))


def em(idx, R, EUMAps, in_model=None, max_iter=200000, eps_abs=1e-7, eps_rel=1e-10, nthreads=0, n_steps=0):
    cp, ct = _i64(idx.class_ptr), _i32(idx.class_tid)
    R = _i32(R)
    ps = np.ascontiguousarray(EUMAps, dtype=np.float64)
    im = None if in_model is None else np.ascontiguousarray(in_model, dtype=np.uint8)
    theta = np.zeros(idx.T)
    steps = np.zeros((n_steps, idx.T)) if n_steps else None
    fd = C.c_double(0)
    it = lib().orc_em(int(idx.T), C.c_int64(idx.C), _p(cp, C.c_int64), _p(ct, C.c_int32), _p(R, C.c_int32), _p(ps, C.c_double),
                      _p(im, C.c_uint8) if im is not None else None, int(max_iter), C.c_double(eps_abs), C.c_double(eps_rel),
                      int(nthreads), _p(theta, C.c_double), C.byref(fd), int(n_steps),
                      _p(steps, C.c_double) if steps is not None else None)
    return theta, int(it), float(fd.value), steps


def em_time(idx, R, EUMAps, iters, nthreads):
    cp, ct = _i64(idx.class_ptr), _i32(idx.class_tid)
    R = _i32(R)
    ps = np.ascontiguousarray(EUMAps, dtype=np.float64)
    theta = np.zeros(idx.T)
    return float(lib().orc_em_time(int(idx.T), C.c_int64(idx.C), _p(cp, C.c_int64), _p(ct, C.c_int32), _p(R, C.c_int32),
                                   _p(ps, C.c_double), int(iters), int(nthreads), _p(theta, C.c_double)))


def finalize(idx, theta, adjEUMA, iEUMA, N):
    cp, ct = _i64(idx.class_ptr), _i32(idx.class_tid)
    th = np.ascontiguousarray(theta, dtype=np.float64)
    adj = np.ascontiguousarray(adjEUMA, dtype=np.float64)
    iE = np.ascontiguousarray(iEUMA, dtype=np.float64)
    ir = np.zeros(idx.T)
    iri = np.zeros(idx.T, dtype=np.int32)
    tpm = np.zeros(idx.T)
    ex = np.zeros(idx.C)
    tot = C.c_int64(0)
    lib().orc_finalize(int(idx.T), C.c_int64(idx.C), _p(cp, C.c_int64), _p(ct, C.c_int32), _p(th, C.c_double),
                       _p(adj, C.c_double), _p(iE, C.c_double), C.c_int64(N), _p(ir, C.c_double), _p(iri, C.c_int32),
                       _p(tpm, C.c_double), _p(ex, C.c_double), C.byref(tot))
    return ir, iri, tpm, ex, int(tot.value)


def quantify(idx, reads, eumacut=0.0, delta=0.0, **em_kw):
    """The whole hot path on the CPU: counts -> Wf/adjEUMA -> sets/EUMAcut -> EM -> outputs."""
    R, F, N = count(idx, reads)
    Wf, adj, ps, iE = prepare(idx, F, N, delta)
    max_sid, cut, CS, TS = components(idx, adj, eumacut)
    in_model = (CS >= 0).astype(np.uint8)
    theta, n_iter, fd, _ = em(idx, R, ps, in_model, **em_kw)
    ir, iri, tpm, ex, tot = finalize(idx, theta, adj, iE, N)
    return dict(ReadCount=R, FraglengthCounts=F, N=N, Wf=Wf, adjEUMA=adj, EUMAps=ps, iEUMA=iE, CS=CS, TS=TS,
                eumacut=cut, max_sid=max_sid, in_model=in_model, fpkm=theta, n_iter=n_iter, final_delta=fd,
                ireadcount=ir, ireadcount_int=iri, tpm=tpm, expected=ex, total_ireadcount=tot,
                loglik=loglik(idx, R, ps, theta, in_model))
