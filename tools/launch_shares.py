"""Developer tool: summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (launches, total, share, mean)
and write a compact copy of the list (id, kernel, grid, block, ns) that is small enough to commit.

    python tools/launch_shares.py gpurun_out/<launches>.csv profiles/<name>_launches_compact.csv > profiles/<name>_launch_shares.txt
"""
import csv
import re
import sys
from collections import defaultdict

src, compact = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows = []
with open(src, newline='') as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {n: i for i, n in enumerate(hdr)}
for r in rd:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']])
    name = re.sub(r'^void ', '', name)
    if len(name) > 70:
        name = name[:70]
    rows.append((int(r[ix['ID']]), name, r[ix['Grid Size']], r[ix['Block Size']], float(r[ix['Metric Value']].replace(',', ''))))
agg = defaultdict(lambda: [0, 0.0])
for _, n, g, b, ns in rows:
    agg[n][0] += 1; agg[n][1] += ns
tot = sum(v[1] for v in agg.values())
print('launches %d, total GPU time %.1f ms (ncu gpu__time_duration: cold-cache, serialised -- compare shares, not absolutes)' % (len(rows), tot / 1e6))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print('%-72s launches %6d  total %12.1f us  share %5.1f%%  avg %9.1f us' % (n, c, t / 1e3, 100 * t / tot, t / c / 1e3))
own = sum(t for n, (c, t) in agg.items() if not n.startswith('at::') and 'elementwise' not in n and 'cub::' not in n and 'nccl' not in n.lower())
print('kernels of this repo: %.1f%% of the GPU time' % (100 * own / tot))
if compact:
    with open(compact, 'w') as f:
        f.write('id,kernel,grid,block,ns\n')
        for i, n, g, b, ns in rows:
            f.write('%d,"%s","%s","%s",%d\n' % (i, n, g, b, ns))
