"""Opcode mix (executed warp instructions) of one kernel of an .ncu-rep:  python profiles/opmix.py REP KERNEL_INDEX"""
import csv, sys, subprocess, collections
rep, kid = sys.argv[1], int(sys.argv[2])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
lo = starts[kid]; hi_ = starts[kid + 1] if kid + 1 < len(starts) else len(rows)
hdr = rows[lo + 1]; data = [r for r in rows[lo + 2:hi_] if len(r) >= len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
c = collections.Counter()
for r in data:
    toks = r[ix['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    c[op.split('.')[0]] += int(r[ix['Instructions Executed']])
tot = sum(c.values())
print(rows[lo][1][:100], 'total', tot)
for k, v in c.most_common(22):
    print(f'  {k:10s} {v:>12d} {100*v/tot:5.1f}%')
