"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every declared symbol, and refuses
to run without a device (there is no CPU fallback)."""
import ctypes as C
import os
import re

from emsar_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "emsar_cuda.h")).read()
    declared = set(re.findall(r"\b(emsar_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for sym in sorted(declared):
        assert hasattr(L, sym), f"{sym} declared in include/emsar_cuda.h but not exported"
    assert declared == set(_lib.SYMBOLS)


def test_open_fails_loudly_without_device(built):
    import torch
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    rc = _lib.lib().emsar_cuda_open(0, C.byref(h))
    assert rc == 1  # EMSAR_ERR_NO_DEVICE
    assert b"no CUDA device" in _lib.lib().emsar_cuda_last_error()
    assert b"no CPU" in _lib.lib().emsar_cuda_strerror(rc) or b"no usable" in _lib.lib().emsar_cuda_strerror(rc)


def test_struct_layouts_match_header(built):
    """ctypes mirrors must have the C sizes (guards against silent ABI drift)."""
    src = r'''
    #include <stdio.h>
    #include "emsar_cuda.h"
    int main(){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(emsar_index_desc), sizeof(emsar_index_info), sizeof(emsar_device_info),
                        sizeof(emsar_solve_opts), sizeof(emsar_solve_out), sizeof(emsar_model_stats)); return 0; }
    '''
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "s")]).split()))
    mine = [C.sizeof(x) for x in (_lib.IndexDesc, _lib.IndexInfo, _lib.DeviceInfo, _lib.SolveOpts, _lib.SolveOut, _lib.ModelStats)]
    assert sizes == mine


def test_locality_order_host_helper(built):
    """emsar_locality_order (host only): a permutation in every mode; identity-like for family blocks; and for a transcriptome under
    random names the automatic mode brings the share of member references that leave an SM-sized range back down."""
    import ctypes as C
    import numpy as np
    from emsar_b200 import synth
    L = _lib.lib()

    def order(idx, mode):
        o = np.zeros(idx.T, dtype=np.int32)
        cp, ct = np.ascontiguousarray(idx.class_ptr, dtype=np.int64), np.ascontiguousarray(idx.class_tid, dtype=np.int32)
        assert L.emsar_locality_order(C.c_int32(idx.T), C.c_int64(idx.C), cp.ctypes.data_as(C.c_void_p), ct.ctypes.data_as(C.c_void_p),
                                      C.c_int32(mode), o.ctypes.data_as(C.c_void_p)) == 0
        return o

    def remote_share(idx, o, nb=16):
        rank = np.empty(idx.T, dtype=np.int64)
        rank[o] = np.arange(idx.T)
        blk = rank * nb // idx.T
        T, cp, ct = idx.T, idx.class_ptr, idx.class_tid
        k = np.diff(cp)[T:]
        mb = blk[ct[T:]]
        first = np.repeat(mb[cp[T:-1] - T], k)
        return float((mb != first).mean())

    plain = synth.make_index_v2(T=20000, n_multi=150000, kmax=40, seed=9, module_cap=600, scatter=False, p_cross=0.0)
    shuf = synth.make_index_v2(T=20000, n_multi=150000, kmax=40, seed=9, module_cap=600, scatter=True, p_cross=0.1, shuffle_tids=True)
    for idx in (plain, shuf):
        for mode in (0, 1, 2, 3):
            assert np.array_equal(np.sort(order(idx, mode)), np.arange(idx.T))
    assert np.array_equal(order(plain, 0), np.arange(plain.T))
    assert remote_share(plain, order(plain, 3)) <= remote_share(plain, order(plain, 0)) + 0.01
    s0, s3 = remote_share(shuf, order(shuf, 0)), remote_share(shuf, order(shuf, 3))
    assert s0 > 0.5 and s3 < 0.25, (s0, s3)
