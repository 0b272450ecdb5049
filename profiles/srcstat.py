"""Stall summary of one kernel of an .ncu-rep (source page, SASS view):  python profiles/srcstat.py REP KERNEL_INDEX [TOP_N]"""
import csv, sys, subprocess
rep, kid = sys.argv[1], int(sys.argv[2])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
lo = starts[kid]; hi_ = starts[kid + 1] if kid + 1 < len(starts) else len(rows)
print(rows[lo][1][:150])
hdr = rows[lo + 1]; data = [r for r in rows[lo + 2:hi_] if len(r) >= len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'sass lines', len(data), 'warp-instr', sum(int(r[ix['Instructions Executed']]) for r in data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print('  ', s, v, '%.1f%%' % (100 * v / max(tot, 1)))
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:n]
for i in sorted(top):
    r = data[i]; st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(i, r[ix['Source']].strip()[:64], r[ix['# Samples']], r[ix['Instructions Executed']], st)
