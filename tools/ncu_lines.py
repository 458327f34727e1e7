#!/usr/bin/env python
"""Warp-stall samples and executed instructions of one profiled kernel per SOURCE LINE: the SASS page of an .ncu-rep
joined with `nvdisasm --print-line-info` of the cubin it was built from (same compiler + flags => same instruction order).
    usage: ncu_lines.py REP CUBIN KERNEL_SYMBOL_SUBSTRING [top N]
    cubin: cuobjdump -xelf all eamm_b200/lib/conv_tc.o"""
import collections, csv, re, subprocess, sys
rep, cubin, sym = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
if sym == "auto":          # conv_tc_kernel<(bool)0, (bool)1> -> ILb0ELb1E
    name = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()[0]
    m = re.search(r"conv_tc_kernel<\(bool\)(\d), \(bool\)(\d)>", name)
    sym = "conv_tc_kernelILb%sELb%sE" % (m.group(1), m.group(2)) if m else "conv_tc_kernel"
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, inside = {}, ("", 0), False
for l in dis:
    if l.startswith(".text."):
        inside = sym in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16) // 16] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
S, X = h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
body = [r for r in rows[hi + 1:] if len(r) > max(S, X) and r[0].startswith("0x")]
if len(body) != len(line_of):
    print("WARNING: %d SASS rows in the report, %d in the cubin function -- different builds?" % (len(body), len(line_of)))
tot = 0
for i, r in enumerate(body):
    ln = line_of.get(i, ("", -1))
    s, x = int(r[S] or 0), int(r[X] or 0)
    agg[ln][0] += s; agg[ln][1] += x; tot += s
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            agg[ln][2][h[c][6:]] += v
srcs = {}
def text(ln):
    f, n = ln
    if f not in srcs:
        try:
            srcs[f] = open(f).read().splitlines()
        except Exception:
            srcs[f] = []
    return srcs[f][n - 1].strip()[:90] if 0 < n <= len(srcs[f]) else ""
totx = sum(v[1] for v in agg.values())
print("samples %d, executed warp instructions %d" % (tot, totx))
print("-- by stall samples")
for ln, (s, x, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% %7d x%-10d %s:%-5d %-28s | %s" % (100.0 * s / max(1, tot), s, x, ln[0].split("/")[-1][:10], ln[1],
                                                     ", ".join("%s %d" % kv for kv in st.most_common(3)), text(ln)))
print("-- by executed warp instructions")
for ln, (s, x, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% x%-10d %s:%-5d | %s" % (100.0 * x / max(1, totx), x, ln[0].split("/")[-1][:10], ln[1], text(ln)))
