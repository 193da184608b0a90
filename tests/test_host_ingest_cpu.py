"""Host ingestion pipeline, CPU only: the threaded BGZF reader and the double-buffered grouper must hand
the device exactly the read groups the single-threaded reader produces (which the golden tests pin against the reference)."""
import os

import numpy as np
import pytest

from emsar_b200 import host, synth
import golden_util as gu


def _same(a, b):
    return (np.array_equal(a.read_ptr, b.read_ptr) and np.array_equal(a.read_tid, b.read_tid) and np.array_equal(a.read_fraglen, b.read_fraglen))


@pytest.fixture(scope="module")
def pe_bam(built, tmp_path_factory):
    """~60K PE fragments, multi-mapping, as a multi-block BAM (several hundred BGZF blocks)."""
    tmp = tmp_path_factory.mktemp("ingest")
    idx = synth.make_index(T=300, n_multi=1500, kmax=12, seed=7, module_cap=40, nF=41, frag_min=60, readlength=50)
    reads = synth.make_reads(idx, 60000, seed=7)
    synth.write_rsh(idx, str(tmp / "in.rsh"))
    synth.write_sam_pe(idx, reads, str(tmp / "in.sam"))
    synth.sam_to_bam(str(tmp / "in.sam"), str(tmp / "in.bam"))
    return tmp


def test_threaded_bgzf_matches_zlib_reader(pe_bam):
    rsh = host.Rsh(str(pe_bam / "in.rsh"))
    ref, rl0 = host.read_alignments(rsh, str(pe_bam / "in.bam"), pe=True, fmt="bam")
    sam, _ = host.read_alignments(rsh, str(pe_bam / "in.sam"), pe=True, fmt="sam")
    assert _same(ref, sam) and len(ref.read_fraglen) > 50000
    for thr in (1, 3, 8):
        for nbuf in (1, 2):
            got, rl = host.read_alignments(rsh, str(pe_bam / "in.bam"), pe=True, fmt="bam", io_threads=thr, nbuf=nbuf, batch_reads=4096)
            assert rl == rl0 and _same(ref, got), (thr, nbuf)
    rsh.close()


def test_golden_bam_through_the_pipeline(built, tmp_path):
    fx = gu.FIXTURES["pe_bam_ssfr"]
    rsh = host.Rsh(gu.materialize(fx["rsh"], tmp_path))
    bam = gu.materialize(fx["aln"], tmp_path)
    a, _ = host.read_alignments(rsh, bam, pe=True, strand=fx["strand"], fmt="bam")
    b, _ = host.read_alignments(rsh, bam, pe=True, strand=fx["strand"], fmt="bam", io_threads=4, nbuf=2, batch_reads=1000)
    assert _same(a, b)
    rsh.close()


def test_corrupt_block_is_an_error(pe_bam, tmp_path):
    raw = bytearray(open(pe_bam / "in.bam", "rb").read())
    raw[len(raw) // 2] ^= 0x5A                      # flip bits inside a compressed block: inflate or the CRC must notice
    bad = tmp_path / "bad.bam"
    bad.write_bytes(bytes(raw))
    rsh = host.Rsh(str(pe_bam / "in.rsh"))
    with pytest.raises(host.HostError):
        host.read_alignments(rsh, str(bad), pe=True, fmt="bam", io_threads=3)
    trunc = tmp_path / "trunc.bam"
    trunc.write_bytes(bytes(raw[: len(raw) // 3]))
    with pytest.raises(host.HostError):
        host.read_alignments(rsh, str(trunc), pe=True, fmt="bam", io_threads=3)
    rsh.close()


def _same_index(a, b):
    return (a.T == b.T and a.C == b.C and a.nF == b.nF and np.array_equal(a.class_ptr, b.class_ptr) and np.array_equal(a.class_tid, b.class_tid)
            and np.array_equal(a.euma, b.euma) and np.array_equal(a.has_node, b.has_node) and a.names == b.names
            and (a.min_fraglength, a.max_fraglength, a.readlength, a.max_t_size, a.frag_min, a.frag_max) ==
            (b.min_fraglength, b.max_fraglength, b.readlength, b.max_t_size, b.frag_min, b.frag_max))


def test_packed_rsh_image_round_trip(built, tmp_path, monkeypatch):
    """§8 f3: the packed image reproduces the parsed text index exactly and is dropped when the text changes."""
    idx = synth.make_index(T=400, n_multi=3000, kmax=15, seed=8, module_cap=50, nF=21, frag_min=40, readlength=25, p_no_node=0.1)
    txt = str(tmp_path / "x.rsh")
    synth.write_rsh(idx, txt)
    a = host.Rsh(txt)
    a.save_packed(txt + ".pack", src=txt)
    b = host.Rsh(txt + ".pack", packed=True, src=txt)
    assert _same_index(a, b) and b.tid(a.names[17]) == a.tid(a.names[17]) and b.tid("no such transcript") == -1
    c = host.Rsh(txt, auto=True)                     # what `emsar -I x.rsh` does
    assert c.from_cache and _same_index(a, c)
    for golden in ("pe.in.rsh", "built.in.rsh"):     # the reference's own text files
        g = gu.materialize(golden, tmp_path)
        ga = host.Rsh(g)
        ga.save_packed(g + ".pack", src=g)
        gb = host.Rsh(g, auto=True)
        assert gb.from_cache and _same_index(ga, gb)
        ga.close(); gb.close()
    # the text changes (same content, new mtime): the image is stale, the text is parsed again and, with the switch on, re-cached
    os.utime(txt, ns=(1, 1))
    with pytest.raises(host.HostError):
        host.Rsh(txt + ".pack", packed=True, src=txt)
    d = host.Rsh(txt, auto=True)
    assert not d.from_cache and _same_index(a, d)
    monkeypatch.setenv("EMSAR_RSH_CACHE", "1")
    e = host.Rsh(txt, auto=True)
    f = host.Rsh(txt, auto=True)
    assert not e.from_cache and f.from_cache and _same_index(a, f)
    # a complete image: the derived arrays (transpose, locality order, reachable classes) travel with it
    T, Cn = a.T, a.C
    k = np.diff(a.class_ptr)[T:]
    cid = np.repeat(np.arange(T, Cn, dtype=np.int32), k)
    tids = a.class_tid[T:]
    o = np.lexsort((cid, tids))
    txm_off = np.concatenate([[0], np.cumsum(np.bincount(tids, minlength=T))]).astype(np.uint32)
    g = host.Rsh(txt)
    g.set_aux(txm_off, cid[o], np.arange(T)[::-1], np.ones(Cn - T, dtype=np.uint8), 7, 42)
    g.save_packed(str(tmp_path / "full.pack"), src=txt)
    h = host.Rsh(str(tmp_path / "full.pack"), packed=True, src=txt)
    ax = h.aux()
    assert _same_index(a, h) and ax is not None and b.aux() is None
    assert np.array_equal(ax["txm_off"], txm_off) and np.array_equal(ax["txm_cid"], cid[o]) and np.array_equal(ax["order"], np.arange(T)[::-1])
    assert ax["insertable"].all() and (ax["n_sets_nocut"], ax["max_set_tids"], ax["nnz_multi"]) == (7, 42, len(cid))
    g.close(); h.close()
    # garbage is rejected
    (tmp_path / "bad.pack").write_bytes(b"EMSARPK2" + b"\0" * 40)
    with pytest.raises(host.HostError):
        host.Rsh(str(tmp_path / "bad.pack"), packed=True)
    for r in (a, b, c, d, e, f):
        r.close()


def test_sam_without_sq_header_is_an_error(pe_bam, tmp_path):
    """samtools 0.1.19 (the reference's reader) aborts with 'missing header'; a silent zero-count run is not acceptable."""
    body = [l for l in open(pe_bam / "in.sam") if not l.startswith("@SQ")]
    p = tmp_path / "nosq.sam"
    p.write_text("".join(body))
    rsh = host.Rsh(str(pe_bam / "in.rsh"))
    with pytest.raises(host.HostError, match="@SQ"):
        host.read_alignments(rsh, str(p), pe=True, fmt="sam")
    rsh.close()


def test_malformed_bam_aux_is_an_error(pe_bam, tmp_path):
    """Aux fields are bounds-checked against the record: an unterminated MD:Z string or an oversized B array must fail, not read on."""
    import struct
    import zlib

    def bam(records):
        out = bytearray(b"BAM\1" + struct.pack("<i", 0) + struct.pack("<i", 1))
        out += struct.pack("<i", 3) + b"T0\0" + struct.pack("<i", 1000)
        for body in records:
            out += struct.pack("<i", len(body)) + body
        res = bytearray()
        for chunk in (bytes(out), b""):
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            cd = co.compress(chunk) + co.flush()
            res += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(cd) + 25) + cd + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
        return bytes(res)

    def rec(aux, l_seq=4):
        core = struct.pack("<iiBBHHHiiii", 0, 10, 3, 255, 4680, 1, 0x43, l_seq, 0, 50, 100)
        return core + b"r0\0" + struct.pack("<I", (l_seq << 4) | 0) + bytes((l_seq + 1) // 2) + bytes(l_seq) + aux

    rsh = host.Rsh(str(pe_bam / "in.rsh"))
    cases = {
        "unterminated_z": rec(b"MDZ4444"),                                    # no NUL before the record ends
        "huge_b_array": rec(b"XBBi" + struct.pack("<i", 1 << 28)),
        "negative_b_count": rec(b"XBBi" + struct.pack("<i", -5)),
        "fields_exceed_record": rec(b"", l_seq=4)[:40],
    }
    for name, body in cases.items():
        p = tmp_path / f"{name}.bam"
        p.write_bytes(bam([body]))
        with pytest.raises(host.HostError):
            host.read_alignments(rsh, str(p), pe=True, fmt="bam")
    rsh.close()


def test_threaded_bowtie_parser_matches_single_thread(built, tmp_path):
    """SE bowtie text parsed by a pool of threads (blocks cut at line breaks, grouping in file order) gives exactly the read groups of the
    single-threaded reader, with and without the strand filter; malformed lines and unknown transcripts are still errors."""
    idx = synth.make_index(T=300, n_multi=1500, kmax=12, seed=17, module_cap=40)
    reads = synth.make_reads(idx, 120000, seed=17)
    synth.write_rsh(idx, str(tmp_path / "in.rsh"))
    p = str(tmp_path / "in.bowtie")
    synth.write_bowtie_se(idx, reads, p)
    lines = open(p).read().splitlines(True)
    for i in range(0, len(lines), 7):                    # some reverse-strand alignments for the strand filter
        f = lines[i].split("\t"); f[1] = "-"; lines[i] = "\t".join(f)
    open(p, "w").writelines(lines)
    rsh = host.Rsh(str(tmp_path / "in.rsh"))
    for strand in ("ns", "ssf", "ssr"):
        a, _ = host.read_alignments(rsh, p, fmt="bowtie", strand=strand, io_threads=0)
        for thr in (2, 5):
            b, _ = host.read_alignments(rsh, p, fmt="bowtie", strand=strand, io_threads=thr, nbuf=2, batch_reads=5000)
            assert _same(a, b), (strand, thr)
    fx = gu.FIXTURES["se"]
    grsh = host.Rsh(gu.materialize(fx["rsh"], tmp_path))
    galn = gu.materialize(fx["aln"], tmp_path)
    a, _ = host.read_alignments(grsh, galn, fmt="bowtie", io_threads=0)
    b, _ = host.read_alignments(grsh, galn, fmt="bowtie", io_threads=4)
    assert _same(a, b)
    bad = str(tmp_path / "bad.bowtie")
    open(bad, "w").writelines(lines[:1000] + ["r\t+\tnot_a_transcript\t0\tAAAA\tAAAA\t0\t\n"] + lines[1000:2000])
    with pytest.raises(host.HostError):
        host.read_alignments(rsh, bad, fmt="bowtie", io_threads=3)
    open(bad, "w").writelines(lines[:1000] + ["garbage line without tabs\n"])
    with pytest.raises(host.HostError):
        host.read_alignments(rsh, bad, fmt="bowtie", io_threads=3)
    rsh.close(); grsh.close()
