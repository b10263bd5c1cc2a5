#!/usr/bin/env python
"""SASS opcode histogram per kernel family of libfastcheb_cuda.so (cuobjdump -sass; runs in the build container).
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "fastchebinterp.jl_b200", "libfastcheb_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = sorted(set(re.findall(r"Function : (\S+)", out)))
dem = dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()))
SPECIAL = ("DMMA", "UBLKCP", "LDGSTS", "SYNCS", "UTMA", "REDG", "ATOM", "BAR", "HMMA", "IMMA", "UTC", "LDCU", "LDC")
fams, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        d = dem[m.group(1)]
        k = re.search(r"(\w+_kernel)", d)
        fam = k.group(1) if k else d[:60]
        cur = fams.setdefault(fam, {"n": 0, "ops": collections.Counter()})
        cur["n"] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["ops"][".".join(op.split(".")[:3]) if op.startswith(SPECIAL) else op.split(".")[0]] += 1
arch = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
print("# SASS opcode histogram per kernel family of libfastcheb_cuda.so (cuobjdump -sass; all instantiations of a family summed)")
print("# cubins in the library:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
print("# tcgen05.mma has no FP64 kind: the FP64 tensor path of sm_100a is mma.sync.m8n8k4.f64 -> DMMA.8x8x4.  UBLKCP = cp.async.bulk")
print("# (TMA engine, 1-D bulk copies), SYNCS = mbarrier, LDGSTS = cp.async, LDCU = uniform constant-bank load (kernel parameters).")
print("# No HMMA / IMMA / UTC*MMA: nothing on this path is sub-FP64 tensor work.\n")
for fam, d in fams.items():
    tot = sum(d["ops"].values())
    print(f"{fam}  ({d['n']} instantiations, {tot} instructions)")
    print("   top: " + ", ".join(f"{k} {v}" for k, v in d["ops"].most_common(12)))
    sp = {k: v for k, v in d["ops"].items() if k.startswith(SPECIAL)}
    if sp:
        print("   tensor / async / sync / constant: " + ", ".join(f"{k} {v}" for k, v in sorted(sp.items())))
    print()
