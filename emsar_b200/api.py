"""Python mirror of the reference's per-sample loop (reference src/emsar_main.c:380-488) over the C ABI.

    ctx   = Context(device)                         # one device, one stream
    index = Index(ctx, synth_or_loaded_index)       # construct_rsh_from_rshfile's product, on the device
    smp   = index.sample()                          # clear_readcounts + calloc FraglengthCounts
    smp.count(read_ptr, read_tid, read_fraglen)     # update_ReadCounts for a batch of read groups
    res   = smp.solve()                             # Wf ... MLE ... iEUMA, numeric part of print_FPKMfinal
    smp.close()

Everything numeric happens in libemsar_cuda.so; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        L.check(L.lib().emsar_cuda_open(int(device), C.byref(self._h)), "emsar_cuda_open")
        self.device = int(device)

    def info(self) -> dict:
        di = L.DeviceInfo()
        L.check(L.lib().emsar_cuda_device_info(self._h, C.byref(di)), "emsar_cuda_device_info")
        return {k: (getattr(di, k).decode() if k == "name" else getattr(di, k)) for k, _ in di._fields_}

    def launches(self) -> int:
        n = C.c_int64(0)
        L.check(L.lib().emsar_cuda_launch_count(self._h, C.byref(n)), "emsar_cuda_launch_count")
        return int(n.value)

    def timer_start(self):
        L.check(L.lib().emsar_cuda_timer_start(self._h), "emsar_cuda_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_double(0)
        L.check(L.lib().emsar_cuda_timer_stop(self._h, C.byref(ms)), "emsar_cuda_timer_stop")
        return float(ms.value)

    def synchronize(self):
        L.check(L.lib().emsar_cuda_synchronize(self._h), "emsar_cuda_synchronize")

    # -- multi-GPU (one sample sharded over the ranks) ---------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        L.check(L.lib().emsar_comm_unique_id(buf), "emsar_comm_unique_id")
        return bytes(buf)

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        L.check(L.lib().emsar_comm_init(self._h, int(rank), int(nranks), buf), "emsar_comm_init")

    def comm_init_torch(self):
        """Convenience: ship the NCCL unique id through an initialised torch.distributed process group."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(self.comm_unique_id()), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        self.comm_init(rank, world, bytes(t.cpu().numpy().tobytes()))

    def comm_info(self) -> dict:
        r, n, pm = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        L.check(L.lib().emsar_comm_info(self._h, C.byref(r), C.byref(n), C.byref(pm)), "emsar_comm_info")
        return {"rank": r.value, "nranks": n.value, "peer_memory": pm.value}

    def comm_destroy(self):
        L.lib().emsar_comm_destroy(self._h)

    def close(self):
        if self._h:
            L.lib().emsar_comm_destroy(self._h)
            L.lib().emsar_cuda_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Index:
    """Packed rsh index on the device. `idx` is any object with the fields of emsar_b200.synth.SynthIndex
    (the host rsh loader produces the same)."""

    def __init__(self, ctx: Context, idx):
        self.ctx = ctx
        self.T, self.C, self.nF = int(idx.T), int(idx.C), int(idx.nF)
        self.max_fraglength = int(idx.max_fraglength)
        cp = _np(idx.class_ptr, np.int64)
        ct = _np(idx.class_tid, np.int32)
        eu = _np(idx.euma, np.int32)
        hn = _np(idx.has_node, np.uint8) if getattr(idx, "has_node", None) is not None else None
        d = L.IndexDesc(self.T, self.C, _ptr(cp), _ptr(ct), self.nF, _ptr(eu), _ptr(hn) if hn is not None else None,
                        int(idx.min_fraglength), int(idx.max_fraglength), int(idx.readlength), int(idx.max_t_size), None)
        self._h = C.c_void_p()
        L.check(L.lib().emsar_index_create(ctx._h, C.byref(d), C.byref(self._h)), "emsar_index_create")

    def info(self) -> dict:
        ii = L.IndexInfo()
        L.check(L.lib().emsar_index_info_get(self._h, C.byref(ii)), "emsar_index_info_get")
        return {k: getattr(ii, k) for k, _ in ii._fields_}

    def sample(self) -> "Sample":
        return Sample(self)

    def close(self):
        if self._h:
            L.lib().emsar_index_destroy(self._h)
            self._h = C.c_void_p()


class Sample:
    def __init__(self, index: Index):
        self.index = index
        self._h = C.c_void_p()
        L.check(L.lib().emsar_sample_begin(index._h, C.byref(self._h)), "emsar_sample_begin")

    # -- counting ---------------------------------------------------------------------------------
    def count(self, read_ptr, read_tid, read_fraglen):
        """Host arrays (numpy, or pinned torch CPU tensors). int64 offsets, int32 tids, int32 fragment lengths."""
        if hasattr(read_ptr, "data_ptr"):   # torch tensors: zero-copy
            n = int(read_ptr.numel()) - 1
            if read_ptr.is_cuda:
                L.check(L.lib().emsar_sample_count_device(self._h, C.c_int64(n), C.c_void_p(read_ptr.data_ptr()),
                                                          C.c_void_p(read_tid.data_ptr()), C.c_void_p(read_fraglen.data_ptr())),
                        "emsar_sample_count_device")
            else:
                L.check(L.lib().emsar_sample_count(self._h, C.c_int64(n), C.c_void_p(read_ptr.data_ptr()),
                                                   C.c_void_p(read_tid.data_ptr()), C.c_void_p(read_fraglen.data_ptr())),
                        "emsar_sample_count")
            return
        rp, rt, fl = _np(read_ptr, np.int64), _np(read_tid, np.int32), _np(read_fraglen, np.int32)
        L.check(L.lib().emsar_sample_count(self._h, C.c_int64(len(rp) - 1), _ptr(rp), _ptr(rt), _ptr(fl)), "emsar_sample_count")

    def count_compact(self, read_len, read_tid, read_fraglen=None, const_fraglen=0):
        """The compact wire form: uint16 lengths, int32 tids, uint16 fragment lengths or None (every group has `const_fraglen`).
        numpy arrays or pinned torch CPU tensors."""
        def ptr(a):
            return C.c_void_p(a.data_ptr()) if hasattr(a, "data_ptr") else _ptr(a)
        n = int(read_len.numel()) if hasattr(read_len, "numel") else len(read_len)
        if not hasattr(read_len, "data_ptr"):
            read_len, read_tid = _np(read_len, np.uint16), _np(read_tid, np.int32)
            read_fraglen = _np(read_fraglen, np.uint16) if read_fraglen is not None else None
        nt = int(read_tid.numel()) if hasattr(read_tid, "numel") else len(read_tid)
        L.check(L.lib().emsar_sample_count_compact(self._h, C.c_int64(n), C.c_int64(nt), ptr(read_len), ptr(read_tid),
                                                   ptr(read_fraglen) if read_fraglen is not None else None, C.c_int32(int(const_fraglen))),
                "emsar_sample_count_compact")

    def set_counts(self, ReadCount, FraglengthCounts):
        R, F = _np(ReadCount, np.int32), _np(FraglengthCounts, np.int32)
        assert len(R) == self.index.C and len(F) == self.index.max_fraglength + 1
        L.check(L.lib().emsar_sample_counts_set(self._h, _ptr(R), _ptr(F)), "emsar_sample_counts_set")

    def counts_allreduce(self):
        """Every rank counted a different slice of the read groups: sum the integer counts over the ranks."""
        L.check(L.lib().emsar_sample_counts_allreduce(self._h), "emsar_sample_counts_allreduce")

    def counts(self):
        R = np.zeros(self.index.C, dtype=np.int32)
        F = np.zeros(self.index.max_fraglength + 1, dtype=np.int32)
        N = C.c_int64(0)
        L.check(L.lib().emsar_sample_counts_get(self._h, _ptr(R), _ptr(F), C.byref(N)), "emsar_sample_counts_get")
        return R, F, int(N.value)

    # -- estimation --------------------------------------------------------------------------------
    @staticmethod
    def _opts(eps_abs=0.0, eps_rel=0.0, max_iter=0, delta=0.0, eumacut=0.0, max_ntid_per_sid=0, in_model=None, sharded=False):
        keep = None
        o = L.SolveOpts(eps_abs, eps_rel, int(max_iter), delta, eumacut, int(max_ntid_per_sid), None, int(bool(sharded)))
        if in_model is not None:
            keep = _np(in_model, np.uint8)
            o.in_model = keep.ctypes.data
        return o, keep

    def _out(self):
        T = self.index.T
        bufs = dict(fpkm=np.zeros(T), efflen=np.zeros(T), ireadcount=np.zeros(T), ireadcount_int=np.zeros(T, dtype=np.int32),
                    tpm=np.zeros(T))
        o = L.SolveOut()
        for k, v in bufs.items():
            setattr(o, k, v.ctypes.data)
        return o, bufs

    @staticmethod
    def _result(o, bufs):
        r = dict(bufs)
        for k in ("n_iter", "final_delta", "loglik", "total_ireadcount", "total_readcount", "eumacut", "max_sid", "em_ms", "prep_ms"):
            r[k] = getattr(o, k)
        return r

    def solve(self, **kw) -> dict:
        opts, keep = self._opts(**kw)
        o, bufs = self._out()
        L.check(L.lib().emsar_sample_solve(self._h, C.byref(opts), C.byref(o)), "emsar_sample_solve")
        return self._result(o, bufs)

    def prepare(self, **kw):
        opts, keep = self._opts(**kw)
        L.check(L.lib().emsar_sample_prepare(self._h, C.byref(opts)), "emsar_sample_prepare")

    def model_stats(self) -> dict:
        st = L.ModelStats()
        L.check(L.lib().emsar_sample_model_stats(self._h, C.byref(st)), "emsar_sample_model_stats")
        return {k: getattr(st, k) for k, _ in st._fields_}

    def em_run(self, max_iter=0, stop_on_conv=True, reset_theta=False):
        it, fd, ms = C.c_int32(0), C.c_double(0), C.c_double(0)
        L.check(L.lib().emsar_sample_em_run(self._h, int(max_iter), int(bool(stop_on_conv)), int(bool(reset_theta)),
                                            C.byref(it), C.byref(fd), C.byref(ms)), "emsar_sample_em_run")
        return int(it.value), float(fd.value), float(ms.value)

    def time_adjeuma(self, reps=5) -> float:
        ms = C.c_double(0)
        L.check(L.lib().emsar_sample_time_adjeuma(self._h, int(reps), C.byref(ms)), "emsar_sample_time_adjeuma")
        return float(ms.value)

    def theta(self):
        th = np.zeros(self.index.T)
        L.check(L.lib().emsar_sample_theta_get(self._h, _ptr(th)), "emsar_sample_theta_get")
        return th

    def theta_randomize(self, seed: int):
        L.check(L.lib().emsar_sample_theta_randomize(self._h, C.c_uint64(int(seed))), "emsar_sample_theta_randomize")

    def finalize(self) -> dict:
        o, bufs = self._out()
        L.check(L.lib().emsar_sample_finalize(self._h, C.byref(o)), "emsar_sample_finalize")
        return self._result(o, bufs)

    def segments(self, want_sets=True):
        Cn = self.index.C
        adj, ex = np.zeros(Cn), np.zeros(Cn)
        cs = np.zeros(Cn, dtype=np.int32) if want_sets else None
        L.check(L.lib().emsar_sample_segments_get(self._h, _ptr(adj), _ptr(ex), _ptr(cs) if want_sets else None), "emsar_sample_segments_get")
        return adj, ex, cs

    def wf(self):
        w = np.zeros(self.index.nF)
        L.check(L.lib().emsar_sample_wf_get(self._h, _ptr(w)), "emsar_sample_wf_get")
        return w

    def close(self):
        if self._h:
            L.lib().emsar_sample_end(self._h)
            self._h = C.c_void_p()


def shard_ranges(weight_prefix, nranks):
    """Host-only helper: nnz-balanced contiguous ranges (emsar_shard_ranges)."""
    wp = _np(weight_prefix, np.int64)
    out = np.zeros(nranks + 1, dtype=np.int64)
    L.check(L.lib().emsar_shard_ranges(C.c_int64(len(wp) - 1), _ptr(wp), int(nranks), _ptr(out)), "emsar_shard_ranges")
    return out
