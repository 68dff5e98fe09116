"""Shares per kernel of ONE step from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv): python scripts/launch_shares.py launches.csv [step_index] [title]
A step starts at a k_pack launch; step_index counts from 0 (1 = the timed resident step of `bench.py --steps 1 --warmup 1`)."""
import collections, csv, io, sys
rows = [l for l in open(sys.argv[1]) if not l.startswith('==')]
r = list(csv.reader(io.StringIO(''.join(rows))))
hdr = r[0]
ni, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
ls = []
for x in r[1:]:
    name = x[ni].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    v = float(x[vi].replace(',', ''))
    us = v / 1000.0 if x[ui] in ('ns', 'nsecond') else v * (1000.0 if x[ui] in ('ms', 'msecond') else 1.0)
    ls.append((name, us))
starts = [i for i, (n, _) in enumerate(ls) if n == 'k_pack'] + [len(ls)]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
step = ls[starts[k]:starts[k + 1]]
if len(sys.argv) > 3:
    print('# ' + sys.argv[3] + '\n')
tot = sum(u for _, u in step)
print(f'{len(step)} launches, {tot / 1000.0:.2f} ms in kernels (steps in the list: {len(starts) - 1}).\n')
agg = collections.OrderedDict()
for n, u in step:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += u
print('| kernel | launches | total us | share |\n|---|---|---|---|')
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| {n} | {c} | {u:.1f} | {u / tot:.3f} |')
