#!/usr/bin/env python
"""Aggregate executed warp-instructions of an ncu source-page CSV by enclosing function of
mdx_kernels.cuh (resolved through nvdisasm's inlined-at line info).
usage: region_profile.py <src.csv> <kernel.sass> <mangled-kernel-substring> <n_envs>"""
import csv, re, sys, collections
src_csv, sass, kname, n_envs = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
cu = open("/root/repo/marl-delivery_b200/csrc/mdx_kernels.cuh").read().split("\n")
# function / section boundaries: device functions, and '// ----' comment banners inside them
marks = []
for i, l in enumerate(cu, 1):
    m = re.match(r"\s*(?:template.*\n)?(?:__device__|__global__).*?\b(\w+)\s*\(", l)
    if m and "(" in l and not l.strip().startswith("//"):
        marks.append((i, m.group(1)))
    m = re.match(r"\s*// -{4,} (.+?) -{4,}", l)
    if m:
        marks.append((i, "  " + m.group(1)[:40]))
def region(ln):
    name = "?"
    for a, n in marks:
        if a <= ln: name = n
        else: break
    return name
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
seq = []; cur = None
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        f, ln, f2, ln2 = m.group(1).split("/")[-1], int(m.group(2)), (m.group(3) or "").split("/")[-1], int(m.group(4) or 0)
        cur = region(ln) if f == "mdx_kernels.cuh" else (region(ln2) + " [lib]" if f2 == "mdx_kernels.cuh" else "lib:" + f)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l): seq.append(cur)
rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[h]
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = collections.Counter(); samp = collections.Counter(); k = 0
for r in rows[h + 1:]:
    if r and r[0] in ("Kernel Name", "Address"): break
    if len(r) > ie:
        key = seq[k] if k < len(seq) else "?"
        agg[key] += int(r[ie]); samp[key] += int(r[isamp]); k += 1
tot = sum(agg.values()); ts = sum(samp.values())
print(f"total warp-instructions {tot}  = {tot/n_envs:.0f} per env-step; stall samples {ts}")
for n, v in agg.most_common():
    print(f"{n:48s} {v/tot:6.1%} inst  {v/n_envs:7.1f}/env   {samp[n]/max(ts,1):6.1%} samples")
