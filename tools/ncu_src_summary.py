"""Summarise the source page of an .ncu-rep (stall samples per SASS instruction, instruction mix).
usage: python tools/ncu_src_summary.py report.ncu-rep [n_top]"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
samp = idx['# Samples']
tot = sum(int(r[samp]) for r in data)
print('kernel', rows[0][1][:80]); print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
    print(f'  {s:26s} {v:8d} {v/tot:.3f}')
for r in sorted(data, key=lambda r: -int(r[samp]))[:ntop]:
    st = sorted(((s, int(r[idx[s]])) for s in stalls if int(r[idx[s]]) > 0), key=lambda x: -x[1])[:2]
    print(f"{int(r[samp]):7d} {int(r[idx['Instructions Executed']]):10d}  {r[1].strip()[:64]:64s} {st}")
c = Counter()
for r in data:
    n = int(r[idx['Instructions Executed']])
    t = r[1].strip().split()
    op = t[1] if t[0].startswith('@') else t[0]
    c[op.split('.')[0]] += n
ti = sum(c.values())
print('total warp instructions', ti)
print('  ' + ', '.join(f'{k} {v/ti:.3f}' for k, v in c.most_common(24)))
