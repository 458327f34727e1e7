#!/usr/bin/env python
"""Hot spots of one kernel from an .ncu-rep (SASS view of `--page source`): instruction mix, executed warp instructions and
stall samples per opcode, and the top stalled instructions with their dominant stall reason.   usage: ncu_hot.py REP [top N]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
S, X, T = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot_s = tot_x = 0
by_op = collections.defaultdict(lambda: [0, 0, 0])
items = []
for r in rows[hi + 1:]:
    if len(r) <= max(S, X):
        continue
    try:
        s, x = int(r[S]), int(r[X])
    except ValueError:
        continue
    op = r[T].split()[0] if r[T].split() else "?"
    if op.startswith("@"):
        op = r[T].split()[1]
    op = op.split(".")[0]
    by_op[op][0] += x; by_op[op][1] += s; by_op[op][2] += 1
    tot_s += s; tot_x += x
    st = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
    items.append((s, x, r[T].strip(), st))
print("total warp instructions %d, samples %d" % (tot_x, tot_s))
print("-- by opcode (executed, share; samples, share; static count)")
for op, (x, s, n) in sorted(by_op.items(), key=lambda kv: -kv[1][1])[:22]:
    print("  %-10s %12d %5.1f%%   %8d %5.1f%%   %4d" % (op, x, 100.0 * x / max(1, tot_x), s, 100.0 * s / max(1, tot_s), n))
print("-- top stalled instructions")
for s, x, t, st in sorted(items, reverse=True)[:top]:
    print("  %7d %5.1f%%  x%-9d %-60s %s" % (s, 100.0 * s / max(1, tot_s), x, t[:60], ", ".join("%s %d" % (n, v) for v, n in st)))
