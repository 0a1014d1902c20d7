"""Correlate an ncu --page source --csv (SASS) dump with source lines via nvdisasm --print-line-info.
usage: ncu_lines.py <sass.csv from ncu> <nvdisasm dump> <kernel symbol substring> [top]"""
import csv, re, sys, collections
src_csv, dis, sym = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: find the .text section of the kernel, walk instructions recording current line
lines = open(dis).read().split("\n")
start = None
for i, l in enumerate(lines):
    if l.startswith("\t.section\t.text.") and sym in l or (l.strip().startswith(".text.") and sym in l):
        start = i
        break
assert start is not None, "kernel text section not found"
cur = None
off2line = {}
fun = {}
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        # inlined at?
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0])
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
stall_tot = collections.Counter()
per_line_stall = collections.defaultdict(collections.Counter)
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    off = int(r[ia], 16) - base
    ln = off2line.get(off)
    s, n = int(r[isamp] or 0), int(r[iinst] or 0)
    agg[ln][0] += s; agg[ln][1] += n; tot_s += s; tot_i += n
    for c in stall_cols:
        v = int(r[c] or 0)
        if v: stall_tot[hdr[c]] += v; per_line_stall[ln][hdr[c]] += v
print("total samples", tot_s, "warp instructions", tot_i)
print("stall totals:", ", ".join(f"{k}={v}" for k, v in stall_tot.most_common(8)))
for ln, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in per_line_stall[ln].most_common(3))
    print(f"{str(ln):40s} samples {s:7d} ({100.0 * s / tot_s:5.1f}%)  inst {n:10d} ({100.0 * n / tot_i:5.1f}%)  {st}")
