"""Summarise an `ncu --page source --csv` dump: instruction mix, stall reasons, hottest SASS regions.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_sass_summary.py src.csv
"""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
S, N = col["# Samples"], col["Instructions Executed"]
tot_s = sum(float(r[S] or 0) for r in body)
tot_n = sum(float(r[N] or 0) for r in body)
print("kernel:", rows[0][1][:100])
print("SASS instructions: %d   warp-instructions executed: %.3e   samples: %d" % (len(body), tot_n, tot_s))
mix = collections.Counter()
smp = collections.Counter()
for r in body:
    op = r[col["Source"]].split()[0] if not r[col["Source"]].strip().startswith("@") else r[col["Source"]].split()[1]
    op = op.split(".")[0]
    mix[op] += float(r[N] or 0)
    smp[op] += float(r[S] or 0)
print("\nopcode            %instr   %samples")
for op, n in mix.most_common(18):
    print("%-16s %7.2f   %7.2f" % (op, 100 * n / tot_n, 100 * smp[op] / tot_s))
cand = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
print("\nstall reason (all samples)   %samples")
tots = {h: sum(float(r[col[h]] or 0) for r in body) for h in cand}
for h, v in sorted(tots.items(), key=lambda kv: -kv[1]):
    if v:
        print("%-28s %6.2f" % (h, 100 * v / tot_s))
# hottest 64-instruction windows
W = 64
win = []
for i in range(0, len(body), W):
    s = sum(float(r[S] or 0) for r in body[i:i + W])
    n = sum(float(r[N] or 0) for r in body[i:i + W])
    win.append((s, n, i))
print("\nhottest %d-instruction windows (samples%%, instr%%, first instruction)" % W)
for s, n, i in sorted(win, reverse=True)[:12]:
    ops = collections.Counter(r[col["Source"]].split()[0].split(".")[0] for r in body[i:i + W])
    print("%6.2f %6.2f  #%5d  %s" % (100 * s / tot_s, 100 * n / tot_n, i, dict(ops.most_common(4))))
