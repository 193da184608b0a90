#!/usr/bin/env python
"""Host ingestion rate: BAM -> read groups, by number of BGZF inflate threads, next to the unmodified
reference binary reading the same file. CPU only (the batches go to a no-op consumer instead of the device).
    python profiles/ingest_bench.py [n_fragments]
"""
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emsar_b200 import host, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
    tmp = tempfile.mkdtemp(prefix="ingest_")
    idx = synth.make_index(T=2000, n_multi=12000, kmax=20, seed=9, module_cap=60, nF=101, frag_min=150, readlength=100)
    reads = synth.make_reads(idx, n, seed=9)
    synth.write_rsh(idx, f"{tmp}/in.rsh")
    t0 = time.time()
    synth.write_sam_pe(idx, reads, f"{tmp}/in.sam")
    synth.sam_to_bam(f"{tmp}/in.sam", f"{tmp}/in.bam")
    n_rec = 2 * int(reads.read_ptr[-1])
    print(f"fixture: {n} PE fragments, {n_rec} BAM records, {os.path.getsize(tmp + '/in.bam') / 1e6:.1f} MB BAM, {os.path.getsize(tmp + '/in.sam') / 1e6:.1f} MB SAM "
          f"(written in {time.time() - t0:.0f} s), host cores {os.cpu_count()}")
    rsh = host.Rsh(f"{tmp}/in.rsh")
    base = None
    for fmt, path, thr in [("sam", "in.sam", 0), ("bam", "in.bam", 0), ("bam", "in.bam", 1), ("bam", "in.bam", 2), ("bam", "in.bam", 4), ("bam", "in.bam", 8)]:
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            r, _ = host.read_alignments(rsh, f"{tmp}/{path}", pe=True, fmt=fmt, io_threads=thr, nbuf=2 if thr else 1)
            best = min(best, time.perf_counter() - t0)
        if base is None:
            base = r
        assert len(r.read_fraglen) == len(base.read_fraglen)
        what = "zlib gz* on the parsing thread" if thr == 0 and fmt == "bam" else (f"{thr} inflate thread(s) + parser" if fmt == "bam" else "text")
        print(f"  {fmt} {what:34s} {best:6.2f} s  {n_rec / best / 1e6:6.2f} M records/s  {len(r.read_fraglen) / best / 1e6:5.2f} M read groups/s")
    ref = os.path.join(ROOT, "oracle", "_ref", "emsar")
    if os.path.exists(ref):
        os.makedirs(f"{tmp}/out", exist_ok=True)
        t0 = time.perf_counter()
        out = subprocess.run([ref, "-n", "1", "-i", "1", "-l", "1", "-P", "-B", "-I", f"{tmp}/in.rsh", f"{tmp}/out", "p", f"{tmp}/in.bam"],
                             capture_output=True, text=True)
        total = time.perf_counter() - t0
        stamps = re.findall(r"^(.*?) :\d\d/\d\d,(\d\d):(\d\d):(\d\d)", out.stdout, re.M)
        sec = [int(h) * 3600 + int(m) * 60 + int(s) for _, h, m, s in stamps]
        read_s = None
        for i, (what, *_ ) in enumerate(stamps):
            if "eading" in what and i + 1 < len(stamps):
                read_s = sec[i + 1] - sec[i]
        print(f"  reference emsar (whole run, -n 1 -i 1): {total:.1f} s; its own stamps put reading the BAM at ~{read_s} s "
              f"(1 s resolution) -> ~{n_rec / max(read_s or total, 1) / 1e6:.2f} M records/s")
    rsh.close()
    # ---- §8 f3: text parse vs packed image of a larger index ----
    big = synth.make_index(T=50000, n_multi=500000, alpha=2.4, kmax=99, seed=10, module_cap=500)
    synth.write_rsh(big, f"{tmp}/big.rsh")
    t0 = time.perf_counter(); a = host.Rsh(f"{tmp}/big.rsh"); t_text = time.perf_counter() - t0
    a.save_packed(f"{tmp}/big.rsh.pack", src=f"{tmp}/big.rsh")
    t0 = time.perf_counter(); b = host.Rsh(f"{tmp}/big.rsh", auto=True); t_pack = time.perf_counter() - t0
    assert b.from_cache and a.C == b.C
    print(f"rsh index T={a.T} C={a.C}: text {os.path.getsize(tmp + '/big.rsh') / 1e6:.0f} MB parsed in {t_text:.2f} s; packed image "
          f"{os.path.getsize(tmp + '/big.rsh.pack') / 1e6:.0f} MB loaded in {t_pack:.2f} s (both include the copy into numpy)")
    a.close(); b.close()


if __name__ == "__main__":
    main()
