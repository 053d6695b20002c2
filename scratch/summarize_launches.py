import csv, collections, re, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
with open(src) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = r['Kernel Name']
    name = name.replace('<unnamed>::', '').replace('cirs::', '').replace('void ', '')
    if name.startswith('gemm_kernel'):
        short = name.split('(')[0]
    elif name.startswith('native::') or name.startswith('at::'):
        short = 'torch:' + re.split(r'[<(]', name)[0].replace('native::', '')
    else:
        short = name.split('(')[0]
    v = float(r['Metric Value'].replace(',', ''))
    v = {'ns': v / 1000.0, 'us': v, 'ms': v * 1000.0}[r['Metric Unit']]
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
out = [title, "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
       "%-78s %8s %12s %8s %9s" % ("kernel", "launches", "total_us", "share", "avg_us")]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-78s %8d %12.1f %7.2f%% %9.2f" % (k[:78], c, t, 100 * t / tot, t / c))
out.append("%-78s %8d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))
open(dst, 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))
