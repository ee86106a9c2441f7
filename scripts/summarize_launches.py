"""Summarise an ncu launch list (gpu__time_duration.sum CSV) of bench.py: one full projection step between two adam launches."""
import csv, collections, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [r['Kernel Name'] for r in rows]
adam = [i for i, n in enumerate(names) if 'adam_noise_kernel' in n]
s, e = adam[1] + 2, adam[2] + 2
def ms(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1e6 if u.startswith('ns') else (v / 1e3 if u.startswith('us') else v)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[s:e]:
    n = r['Kernel Name'].replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
    n = re.sub(r'\(.*', '', n)
    n = re.sub(r'^void ', '', n)
    if 'conv_tc_kernel' in n or 'conv_tc2_kernel' in n or 'conv_halo' in n: n = re.sub(r'^.*tc::', '', n)
    else: n = re.sub(r'<.*', '', n); n = n.split('::')[-1] if 'mgf::' in n else 'torch: ' + n[-40:]
    agg[n][0] += 1; agg[n][1] += ms(r)
tot = sum(v[1] for v in agg.values())
print("one step: %d launches, %.2f ms summed kernel time" % (e - s, tot))
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
tt = [0, 0.0]
for n, (c, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if n.startswith('torch:'): tt[0] += c; tt[1] += m; continue
    print("| `%s` | %d | %.3f | %.1f%% |" % (n, c, m, 100 * m / tot))
print("| PyTorch kernels (buffer fills / copies, loss bookkeeping) | %d | %.3f | %.1f%% |" % (tt[0], tt[1], 100 * tt[1] / tot))
