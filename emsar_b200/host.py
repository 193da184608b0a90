"""ctypes binding of libemsar_host.so: the C rsh loader, alignment readers and writers (emsar_b200/host/)."""
from __future__ import annotations

import ctypes as C
import os
import types

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_PKG, "libemsar_host.so")
_lib = None
ERRLEN = 512


class HostError(RuntimeError):
    pass


class _Rsh(C.Structure):
    _fields_ = [("T", C.c_int32), ("C", C.c_int64), ("class_ptr", C.POINTER(C.c_int64)), ("class_tid", C.POINTER(C.c_int32)),
                ("nF", C.c_int32), ("euma", C.POINTER(C.c_int32)), ("has_node", C.POINTER(C.c_uint8)),
                ("min_fraglength", C.c_int32), ("max_fraglength", C.c_int32), ("readlength", C.c_int32), ("max_t_size", C.c_int32),
                ("frag_min", C.c_int32), ("frag_max", C.c_int32), ("names", C.POINTER(C.c_char_p)),
                ("name_slots", C.POINTER(C.c_uint32)), ("name_mask", C.c_uint32),
                ("has_aux", C.c_int), ("aux_owned", C.c_int), ("aux_nnz_multi", C.c_int64), ("aux_txm_off", C.POINTER(C.c_uint32)),
                ("aux_txm_cid", C.POINTER(C.c_int32)), ("aux_order", C.POINTER(C.c_int32)), ("aux_insertable", C.POINTER(C.c_uint8)),
                ("aux_n_sets_nocut", C.c_int32), ("aux_max_set_tids", C.c_int32)]


class _ReaderOpts(C.Structure):
    _fields_ = [("pe", C.c_int), ("strand", C.c_char), ("max_repeat", C.c_int), ("format", C.c_char), ("batch_reads", C.c_int64),
                ("io_threads", C.c_int), ("nbuf", C.c_int), ("buf_alloc", C.c_void_p), ("buf_free", C.c_void_p), ("hook_user", C.c_void_p)]


_BATCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise HostError(f"{_SO} is missing: run `python -m emsar_b200.build`")
        L = C.CDLL(_SO)
        L.emsar_rsh_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(_Rsh)), C.c_char_p]
        L.emsar_rsh_load_packed.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.POINTER(_Rsh)), C.c_char_p]
        L.emsar_rsh_save_packed.argtypes = [C.POINTER(_Rsh), C.c_char_p, C.c_char_p, C.c_char_p]
        L.emsar_rsh_load_auto.argtypes = [C.c_char_p, C.POINTER(C.POINTER(_Rsh)), C.POINTER(C.c_int), C.c_char_p]
        L.emsar_rsh_free.argtypes = [C.POINTER(_Rsh)]
        L.emsar_rsh_tid.argtypes = [C.POINTER(_Rsh), C.c_char_p]
        L.emsar_rsh_write.argtypes = [C.POINTER(_Rsh), C.c_int, C.c_char_p, C.c_char_p]
        L.emsar_read_alignments.argtypes = [C.POINTER(_Rsh), C.c_char_p, C.POINTER(_ReaderOpts), C.POINTER(C.c_int), _BATCH_FN, C.c_void_p, C.c_char_p]
        _lib = L
    return _lib


class Rsh:
    """A loaded `.rsh` index. Exposes the same fields as emsar_b200.synth.SynthIndex (numpy views are copies)."""

    def __init__(self, path, packed=False, src=None, auto=False):
        """Text `.rsh` (default), a packed image (`packed=True`; `src` = the text file it must match), or `auto`: what the
        command line does for -I (the fresh `<path>.pack` if there is one, else the text)."""
        self._p = C.POINTER(_Rsh)()
        self.from_cache = bool(packed)
        err = C.create_string_buffer(ERRLEN)
        if packed:
            rc = lib().emsar_rsh_load_packed(os.fsencode(path), os.fsencode(src) if src else None, C.byref(self._p), err)
            if rc == 2:
                raise HostError("stale packed image")
        elif auto:
            fc = C.c_int(0)
            rc = lib().emsar_rsh_load_auto(os.fsencode(path), C.byref(self._p), C.byref(fc), err)
            self.from_cache = bool(fc.value)
        else:
            rc = lib().emsar_rsh_load(os.fsencode(path), C.byref(self._p), err)
        if rc:
            raise HostError(err.value.decode())
        r = self._p.contents
        self.T, self.C, self.nF = int(r.T), int(r.C), int(r.nF)
        self.class_ptr = np.ctypeslib.as_array(r.class_ptr, shape=(self.C + 1,)).copy()
        nnz = int(self.class_ptr[-1])
        self.class_tid = np.ctypeslib.as_array(r.class_tid, shape=(max(nnz, 1),))[:nnz].copy()
        self.euma = np.ctypeslib.as_array(r.euma, shape=(self.C, self.nF)).copy()
        self.has_node = np.ctypeslib.as_array(r.has_node, shape=(self.C,)).copy()
        self.min_fraglength, self.max_fraglength = int(r.min_fraglength), int(r.max_fraglength)
        self.readlength, self.max_t_size = int(r.readlength), int(r.max_t_size)
        self.frag_min, self.frag_max = int(r.frag_min), int(r.frag_max)
        self.names = [r.names[t].decode() for t in range(self.T)]
        self.has_aux = bool(r.has_aux)
        self._aux_keep = None

    def aux(self):
        """The derived arrays a complete packed image carries (None for a text index / an image without them)."""
        r = self._p.contents
        if not r.has_aux:
            return None
        n = int(r.aux_nnz_multi)
        return dict(nnz_multi=n, txm_off=np.ctypeslib.as_array(r.aux_txm_off, shape=(self.T + 1,)).copy(),
                    txm_cid=np.ctypeslib.as_array(r.aux_txm_cid, shape=(max(n, 1),))[:n].copy(),
                    order=np.ctypeslib.as_array(r.aux_order, shape=(self.T,)).copy(),
                    insertable=np.ctypeslib.as_array(r.aux_insertable, shape=(max(self.C - self.T, 1),))[:self.C - self.T].copy(),
                    n_sets_nocut=int(r.aux_n_sets_nocut), max_set_tids=int(r.aux_max_set_tids))

    def set_aux(self, txm_off, txm_cid, order, insertable, n_sets_nocut, max_set_tids):
        """Attach derived arrays (as emsar_index_aux_get hands them out) so that save_packed writes a complete image."""
        keep = [np.ascontiguousarray(txm_off, dtype=np.uint32), np.ascontiguousarray(txm_cid, dtype=np.int32),
                np.ascontiguousarray(order, dtype=np.int32), np.ascontiguousarray(insertable, dtype=np.uint8)]
        self._aux_keep = keep
        r = self._p.contents
        r.has_aux, r.aux_owned, r.aux_nnz_multi = 1, 0, len(keep[1])
        r.aux_txm_off = keep[0].ctypes.data_as(C.POINTER(C.c_uint32)); r.aux_txm_cid = keep[1].ctypes.data_as(C.POINTER(C.c_int32))
        r.aux_order = keep[2].ctypes.data_as(C.POINTER(C.c_int32)); r.aux_insertable = keep[3].ctypes.data_as(C.POINTER(C.c_uint8))
        r.aux_n_sets_nocut, r.aux_max_set_tids = int(n_sets_nocut), int(max_set_tids)
        self.has_aux = True

    def tid(self, name: str) -> int:
        return int(lib().emsar_rsh_tid(self._p, name.encode()))

    def write(self, path, pe=False):
        err = C.create_string_buffer(ERRLEN)
        if lib().emsar_rsh_write(self._p, int(bool(pe)), os.fsencode(path), err):
            raise HostError(err.value.decode())

    def save_packed(self, path, src=None):
        err = C.create_string_buffer(ERRLEN)
        if lib().emsar_rsh_save_packed(self._p, os.fsencode(path), os.fsencode(src) if src else None, err):
            raise HostError(err.value.decode())

    def close(self):
        if self._p:
            lib().emsar_rsh_free(self._p)
            self._p = C.POINTER(_Rsh)()


def read_alignments(rsh: Rsh, path, pe=False, strand="ns", max_repeat=100, fmt="bowtie", batch_reads=0, io_threads=0, nbuf=1):
    """Runs the C reader over an alignment file and returns the concatenated read groups
    (read_ptr int64[n+1], read_tid int32[], read_fraglen int32[n]) plus the PE read length seen."""
    st = {"ns": b"\0", "ssf": b"+", "ssr": b"-", "ssfr": b"+", "ssrf": b"-"}[strand]
    f = {"bowtie": b"\0", "sam": b"s", "bam": b"b"}[fmt]
    o = _ReaderOpts(int(bool(pe)), st, int(max_repeat), f, int(batch_reads), int(io_threads), int(nbuf), None, None, None)
    ptrs, tids, fls = [np.zeros(1, dtype=np.int64)], [], []
    base = [0]

    def cb(user, n, ptr, tid, fl):
        p = np.ctypeslib.as_array(ptr, shape=(n + 1,)).copy()
        nt = int(p[-1])
        tids.append(np.ctypeslib.as_array(tid, shape=(max(nt, 1),))[:nt].copy())
        fls.append(np.ctypeslib.as_array(fl, shape=(n,)).copy())
        ptrs.append(p[1:] + base[0])
        base[0] += nt
        return 0

    rl = C.c_int(rsh.readlength)
    err = C.create_string_buffer(ERRLEN)
    fn = _BATCH_FN(cb)
    if lib().emsar_read_alignments(rsh._p, os.fsencode(path), C.byref(o), C.byref(rl), fn, None, err):
        raise HostError(err.value.decode())
    reads = types.SimpleNamespace(read_ptr=np.concatenate(ptrs), read_tid=np.concatenate(tids) if tids else np.zeros(0, np.int32),
                                  read_fraglen=np.concatenate(fls) if fls else np.zeros(0, np.int32))
    return reads, int(rl.value)
