"""Summarise an ncu launch list that carries gpu__time_duration.sum + dram__bytes_read.sum + dram__bytes_write.sum per launch
(one eager projection step of bench.py between two adam launches): per kernel family launches, ms, DRAM GB read / written, GB/s
against MEASURED_PEAKS.json's HBM peak.  Also writes the DRAM traffic of the convolution launches of the step as JSON (argv[2]), which
bench.py reports as `roofline.traffic` when the live launch count still matches.
    python scripts/summarize_traffic.py gpurun_out/traffic_r02e.csv profiles/r02e_conv_dram_traffic.json [commit]"""
import csv, collections, json, os, re, sys

path = sys.argv[1]
out_json = sys.argv[2] if len(sys.argv) > 2 else None
commit = sys.argv[3] if len(sys.argv) > 3 else ""
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]

with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
launches = collections.OrderedDict()            # ID -> {name, grid, metrics}
for r in csv.DictReader(lines):
    d = launches.setdefault(int(r['ID']), {"name": r['Kernel Name'], "grid": r['Grid Size'], "m": {}})
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    if r['Metric Name'].startswith('gpu__time_duration'):
        v = v / 1e6 if u.startswith('ns') else (v / 1e3 if u.startswith('us') else (v * 1e3 if u.startswith('s') else v))      # -> ms
    else:
        v = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)                                            # -> bytes
    d["m"][r['Metric Name']] = v
rows = [launches[k] for k in sorted(launches)]
adam = [i for i, r in enumerate(rows) if 'adam_noise_kernel' in r["name"]]
s, e = adam[1] + 2, adam[2] + 2              # same window as summarize_launches.py: one full step


def fam(n):
    n = n.replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
    n = re.sub(r'\(.*', '', n)
    n = re.sub(r'^void ', '', n)
    if 'conv_tc_kernel' in n or 'conv_tc2_kernel' in n or 'conv_halo' in n:
        return re.sub(r'^.*tc::', '', n)
    n = re.sub(r'<.*', '', n)
    return n.split('::')[-1] if 'mgf::' in n else 'torch: ' + n[-40:]


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
conv = [0, 0.0, 0.0, 0.0]
for r in rows[s:e]:
    m = r["m"]
    t, rd, wr = m.get('gpu__time_duration.sum', 0.0), m.get('dram__bytes_read.sum', 0.0), m.get('dram__bytes_write.sum', 0.0)
    a = agg[fam(r["name"])]
    a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr
    if 'conv_tc' in r["name"] or 'conv_halo' in r["name"]:
        conv[0] += 1; conv[1] += t; conv[2] += rd; conv[3] += wr
tot = sum(v[1] for v in agg.values())
print("one step: %d launches, %.2f ms summed kernel time, %.2f GB DRAM read + %.2f GB written" % (
    e - s, tot, sum(v[2] for v in agg.values()) / 1e9, sum(v[3] for v in agg.values()) / 1e9))
print("HBM peak %.1f GB/s (MEASURED_PEAKS.json); DRAM GB/s = (read + written) / summed launch time of the family\n" % peak)
print("| kernel | launches | ms | share | DRAM read GB | written GB | DRAM GB/s | of HBM peak |\n|---|---:|---:|---:|---:|---:|---:|---:|")
tt = [0, 0.0, 0.0, 0.0]
for n, (c, m, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if n.startswith('torch:'):
        tt[0] += c; tt[1] += m; tt[2] += rd; tt[3] += wr
        continue
    gbs = (rd + wr) / (m * 1e-3) / 1e9 if m > 0 else 0.0
    print("| `%s` | %d | %.3f | %.1f%% | %.3f | %.3f | %.0f | %.2f |" % (n, c, m, 100 * m / tot, rd / 1e9, wr / 1e9, gbs, gbs / peak))
print("| PyTorch kernels (buffer fills / copies, loss bookkeeping) | %d | %.3f | %.1f%% | %.3f | %.3f | | |" % (tt[0], tt[1], 100 * tt[1] / tot, tt[2] / 1e9, tt[3] / 1e9))
print("\nconvolution launches (conv_tc / conv_tc2 / conv_halo): %d, %.3f ms, DRAM %.2f GB read + %.2f GB written = %.2f GB" % (
    conv[0], conv[1], conv[2] / 1e9, conv[3] / 1e9, (conv[2] + conv[3]) / 1e9))
if out_json:
    json.dump({"conv_launches_per_step": conv[0], "dram_bytes_read": conv[2], "dram_bytes_write": conv[3], "kernel_time_ns": conv[1] * 1e6,
               "commit": commit,
               "source": "%s (ncu dram__bytes_read.sum + dram__bytes_write.sum of the %d conv launches of one eager step, "
                         "--clock-control none, B200, 8 images at 1024^2, commit %s)" % (os.path.basename(path), conv[0], commit)},
              open(out_json, "w"), indent=1)
