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
