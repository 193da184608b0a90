"""The boundary proven, not sketched: oracle/_ref/emsar_cuda is the REFERENCE's own program (its sources, compiled where they lie by
oracle/Makefile:ref_cuda) with the three seams of INTEGRATION.md bound to libemsar_cuda.so by oracle/ref_cuda_shim.c - its hook
pointers (emsar.h:219-221) feed emsar_sample_count, its ReadCount[] comes from emsar_sample_counts_get, its FPKM[] from
emsar_sample_solve; option parsing, readers, Wf, adjEUMA, sets, EUMAps, iEUMA and the writers stay the reference's code. Its output files
must equal what the unmodified reference wrote for the same fixture (integers and deterministic columns exactly, estimates within the
tolerance policy), and what this repository's own `emsar` writes."""
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCHED = os.path.join(ROOT, "oracle", "_ref", "emsar_cuda")
OURS = os.path.join(ROOT, "emsar_b200", "bin", "emsar")


@pytest.mark.parametrize("name", ["se", "pe", "bigmod"])
def test_reference_main_bound_to_libemsar_cuda(built, name, tmp_path):
    if not os.path.exists(PATCHED):
        pytest.skip("oracle/_ref/emsar_cuda was not built (needs the reference sources at build time)")
    fx = gu.FIXTURES[name]
    rsh, aln = gu.materialize(fx["rsh"], tmp_path), gu.materialize(fx["aln"], tmp_path)
    flags = ["-q", "-g", "-n", "2"] + (["-P"] if fx["pe"] else []) + ({"sam": ["-S"], "bam": ["-B"], "bowtie": []}[fx["fmt"]])
    out = str(tmp_path / "patched")
    r = subprocess.run([PATCHED] + flags + ["-I", rsh, out, "p", aln], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "emsar_cuda: EM finished" in r.stdout
    seg_ref, fl_ref, g = gu.read_segments(fx["out"]), gu.read_fraglength(fx["out"]), gu.read_fpkm(fx["out"])
    assert gu.parse_out_file(out + "/p.0.fraglength_effect") == fl_ref["text"]
    seg = gu.parse_out_file(out + "/p.0.segments")
    assert [x[:6] for x in seg] == [x[:6] for x in seg_ref["text"]]          # incl. Readcount: counted on the device through the reference's hooks
    mine = gu.parse_out_file(out + "/p.0.fpkm")
    assert [x[0] for x in mine] == g["names"] and [x[3] for x in mine] == [f"{e:f}" for e in g["efflen"]]
    N = int(fl_ref["counts"].sum())
    fp = np.array([float(x[1]) for x in mine])
    tol = gu.fpkm_tolerance(g, g["efflen"], max(N, 1), fx["rounds"], from_files=True)
    ident = seg_ref["adjEUMA"][:len(fp)] > 0
    assert (np.abs(fp - g["fpkm"])[ident] <= tol[ident]).all()
    # and the same numbers as this repository's own command line (both print the device's FPKM with 6 decimals; the reference's
    # print_FPKMfinal averages the two identical rounds)
    out2 = str(tmp_path / "ours")
    r2 = subprocess.run([OURS] + [f for f in flags if f not in ("-n", "2")] + ["-I", rsh, out2, "p", aln], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    ours = gu.parse_out_file(out2 + "/p.0.fpkm")
    for col in (1, 4, 6):
        a, b = np.array([float(x[col]) for x in mine]), np.array([float(x[col]) for x in ours])
        assert np.allclose(a, b, rtol=0, atol=2e-6), col
    assert [x[5] for x in mine] == [x[5] for x in ours]                        # iReadcount.int
