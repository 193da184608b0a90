"""Index construction: the device builder (csrc/build.cu, `emsar-build --device 0`) next to the host builder (host/build_index.c, -p 16) and the
unmodified reference `emsar-build` (oracle/_ref, where present and within its time limit) on generated transcriptomes; outputs must be the
same bytes. usage: python profiles/build_bench.py [se_T] [pe_T]
EMSAR_BENCH_REF=1 adds the reference (minutes of CPU time: run it on a CPU box with EMSAR_BENCH_NO_DEVICE=1, the signatures are comparable)."""
import hashlib
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINE = os.path.join(ROOT, "emsar_b200", "bin", "emsar-build")
REF = os.path.join(ROOT, "oracle", "_ref", "emsar-build")
TMP = os.environ.get("TMPDIR", "/tmp") + "/emsar_build_bench"


def make_fasta(path, T, seed=7):
    """gene families of 1-6 isoforms over a shared exon pool (80-400 bases), 3 % of the genes reuse exons of a paralog"""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    n_exons = max(64, T * 3)
    exons = [letters[rng.integers(0, 4, size=int(rng.integers(80, 400)))].tobytes() for _ in range(n_exons)]
    t = 0
    bases = 0
    with open(path, "wb") as f:
        g = 0
        while t < T:
            base = (7 * g) % (n_exons - 12)
            pool = [exons[base + j] for j in range(10)]
            if rng.random() < 0.03:
                other = int(rng.integers(0, n_exons - 4))
                pool[3:6] = exons[other:other + 3]
            for iso in range(int(rng.integers(1, 7))):
                keep = [e for e in pool if rng.random() < 0.7] or pool[:1]
                s = b"".join(keep)
                f.write(b">T%d\n" % t + s + b"\n")
                bases += len(s)
                t += 1
                if t >= T:
                    break
            g += 1
    return bases


def run(tool, args, outdir, limit):
    t0 = time.perf_counter()
    try:
        r = subprocess.run([tool] + args + [outdir, "x"], capture_output=True, text=True, timeout=limit, env=dict(os.environ, EMSAR_BUILD_TIMING="1"))
    except subprocess.TimeoutExpired:
        return None, None
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        print("FAILED", tool, args, r.stdout[-500:], r.stderr[-500:])
        return None, None
    data = open(os.path.join(outdir, "x.rsh"), "rb").read()
    for line in r.stderr.splitlines():
        if line.startswith("build timing"):
            print("   ", os.path.basename(tool), " ".join(args[:3]), "|", line, flush=True)
    return dt, (hashlib.sha256(data).hexdigest()[:16], data.count(b"\n"))


def case(label, fa, flags, rl, ref_limit):
    res = {}
    for tag, tool, extra, limit in (("device", MINE, ["--device", "0"], 1200), ("host_p16", MINE, ["-p", "16"], 1200), ("reference", REF, [], ref_limit)):
        if not os.path.exists(tool) or limit <= 0 or (tag == "reference" and not os.environ.get("EMSAR_BENCH_REF")):
            continue
        if tag == "device" and os.environ.get("EMSAR_BENCH_NO_DEVICE"):
            continue
        dt, sig = run(tool, ["-q"] + extra + flags + [fa, rl], os.path.join(TMP, label + "_" + tag), limit)
        res[tag] = (dt, sig)
        print(f"{label:10s} {tag:10s} {'%.2f s' % dt if dt else 'skipped / over the time limit'}  {sig}", flush=True)
    sigs = {v[1] for v in res.values() if v[1]}
    print(f"{label:10s} identical outputs: {len(sigs) == 1}  device vs host speed-up: "
          f"{res['host_p16'][0] / res['device'][0] if res.get('device', (None,))[0] and res.get('host_p16', (None,))[0] else float('nan'):.1f}x", flush=True)


if __name__ == "__main__":
    os.makedirs(TMP, exist_ok=True)
    se_T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    pe_T = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    fa = os.path.join(TMP, "se.fa")
    print("single-end transcriptome:", se_T, "transcripts,", make_fasta(fa, se_T), "bases", flush=True)
    case("se_ns_50", fa, [], "50", 600)
    case("se_ssf_75", fa, ["-s", "ssf"], "75", 0)
    fa2 = os.path.join(TMP, "pe.fa")
    print("paired-end transcriptome:", pe_T, "transcripts,", make_fasta(fa2, pe_T, seed=9), "bases", flush=True)
    case("pe_ns", fa2, ["-P", "-f", "200", "-F", "260"], "50", 0)
    case("pe_ssfr", fa2, ["-P", "-s", "ssfr", "-f", "200", "-F", "260"], "50", 0)
