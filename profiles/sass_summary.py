#!/usr/bin/env python
"""Instruction mix and memory-instruction evidence of the hot kernels from the built objects (cuobjdump -sass), for profiles/.
usage: python profiles/sass_summary.py > profiles/r2_sass.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "emsar_b200", "csrc", "_obj")
KERNELS = [("em_psum.o", "k_em_psumILi0E"), ("prep.o", "k_adjeuma_stream"), ("count.o", "k_count"), ("em.o", "k_em_persistentILi3E"),
           ("build.o", "k_window_hash"), ("build.o", "k_occ_peILb0E"), ("build.o", "k_run_heads")]

for obj, pat in KERNELS:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    cur, body = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
            body[cur].append(line)
    for fn, lines in body.items():
        if pat not in fn:
            continue
        ops = collections.Counter()
        for l in lines:
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", l)
            if m:
                ops[m.group(1)] += 1
        fam = collections.Counter()
        for op, n in ops.items():
            fam[op.split(".")[0]] += n
        print(f"== {fn}: {sum(ops.values())} SASS instructions")
        print("   by opcode family:", ", ".join(f"{k} {v}" for k, v in fam.most_common(18)))
        mem = {k: v for k, v in ops.items() if k.split(".")[0] in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDGSTS", "ATOMS", "ATOMG", "RED", "LDGDEPBAR", "DEPBAR", "UBLKCP", "SYNCS", "MUFU", "DFMA", "DADD", "DMUL", "SHFL", "BAR", "WARPSYNC")}
        print("   memory / fp64 / sync:", ", ".join(f"{k} {v}" for k, v in sorted(mem.items(), key=lambda x: -x[1])))
        print()
