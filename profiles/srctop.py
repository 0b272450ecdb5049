"""Top stall lines of an `ncu --page source --csv` export (SASS view).   python profiles/srctop.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
si = hdr.index("Warp Stall Sampling (All Samples)")
src = hdr.index("Source")
stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_")]
body = []
for r in rows[h + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break  # a second table (same kernel captured twice) follows
    if len(r) > si:
        body.append(r)
tot = sum(float(r[si] or 0) for r in body)
print("total samples", tot)
agg = {}
for c_i, c in stall_cols:
    agg[c] = sum(float(r[c_i] or 0) for r in body)
print({k: round(v / max(tot, 1), 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -float(body[i][si] or 0))[:n]
for i in sorted(order):
    r = body[i]
    st = sorted(((float(r[c_i] or 0), c) for c_i, c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {float(r[si]):7.0f} {100*float(r[si])/tot:5.1f}%  {r[src][:90]:90s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
