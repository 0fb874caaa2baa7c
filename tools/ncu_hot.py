"""Top stall sites of one kernel in an .ncu-rep (source page, SASS view): python tools/ncu_hot.py REP KERNEL_INDEX [N]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
s = starts[kid]
e = starts[kid + 1] if kid + 1 < len(starts) else len(rows)
hdr = rows[s + 1]
si, ss, ie = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
tot = sum(int(r[ss]) for r in body)
print(rows[s][1][:120], '| total samples', tot, '| SASS instructions', len(body))
agg = {}
for r in body:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print('stall totals:', sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ss]))[:n]
for i in sorted(idx):
    r = body[i]
    top = sorted(((int(r[j]), hdr[j][6:]) for j in stalls if int(r[j])), reverse=True)[:3]
    print(f'{i:5d} {int(r[ss]):6d} {100*int(r[ss])/tot:5.1f}% exec={r[ie]:>8s} {r[si].strip()[:70]:70s} {top}')
