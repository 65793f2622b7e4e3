#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump with `nvdisasm -g -c` line info and aggregate executed
warp-instructions and stall samples per CUDA source line.
usage: line_profile.py <src.csv> <kernel.sass> <mangled-kernel-substring> [top_n]"""
import csv, re, sys, collections
src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
# --- sass with line info: sequence of (line_no) per instruction for the kernel section
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
seq = []
cur = None
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or (l.startswith(".text.") ):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(cur)
rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ie, isamp, ithr = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
data = []
for r in rows[h + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break  # first captured launch only
    if len(r) > ie:
        data.append(r)
print("sass insts (nvdisasm):", len(seq), " ncu rows:", len(data), file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0])
for k, r in enumerate(data):
    key = seq[k] if k < len(seq) else None
    a = agg[key]
    a[0] += int(r[ie]); a[1] += int(r[isamp]); a[2] += int(r[ithr])
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print(f"total warp-inst {tot}  samples {tots}")
srcs = {}
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    txt = ""
    if key:
        f = key[0]
        if f not in srcs:
            import glob
            c = glob.glob(f"/root/repo/marl-delivery_b200/csrc/{f}")
            srcs[f] = open(c[0]).read().split("\n") if c else []
        if srcs[f] and key[1] - 1 < len(srcs[f]):
            txt = srcs[f][key[1] - 1].strip()[:90]
    print(f"{str(key):34s} inst {a[0]:11d} {a[0]/tot:6.1%}  samp {a[1]/max(tots,1):6.1%} thr/inst {a[2]/max(a[0],1):5.1f} | {txt}")
