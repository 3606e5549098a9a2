"""Per-SASS-instruction shared-memory wavefronts of a captured kernel: which loads / stores
pay bank conflicts?  Usage: python tools/ncu_smem_lines.py prof.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
start = blocks[0]
end = blocks[1] if len(blocks) > 1 else len(rows)
h = rows[start + 1]
print('columns:', [c for c in h if 'hared' in c or 'avefront' in c or 'onflict' in c])
col = {n: i for i, n in enumerate(h)}
wf = [c for c in h if 'L1 Wavefronts Shared' in c]
key_ex = next((c for c in wf if 'Excessive' in c), None)
key_all = next((c for c in wf if c.strip() == 'L1 Wavefronts Shared'), wf[0] if wf else None)
key_ideal = next((c for c in wf if 'Ideal' in c), None)
print('using', key_all, key_ideal, key_ex)
agg = []
tot = [0, 0]
for r in rows[start + 2:end]:
    if len(r) < len(h):
        continue
    def num(k):
        try:
            return int(float(r[col[k]] or 0)) if k else 0
        except ValueError:
            return 0
    a, i = num(key_all), num(key_ideal)
    if a:
        tot[0] += a
        tot[1] += i
        agg.append((a - i, a, i, r[col['Address']][-5:], r[col['Source']][:80]))
print('total shared wavefronts %d, ideal %d, excess %d' % (tot[0], tot[1], tot[0] - tot[1]))
agg.sort(key=lambda t: -t[0])
for x in agg[:ntop]:
    print('excess %9d  total %9d  ideal %9d  %s  %s' % x)
