"""per-stage table of an ncu launch list of the HBM wavefront: usage hbm_stage_table.py <launches.csv>"""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; ki = h.index('Kernel Name'); mi = h.index('Metric Name'); vi = h.index('Metric Value'); idi = h.index('ID')
per = collections.defaultdict(dict)
for r in rows[1:]:
    try: per[(r[idi], r[ki])][r[mi]] = float(r[vi].replace(',', ''))
    except ValueError: pass
agg = collections.defaultdict(collections.Counter); cnt = collections.Counter()
names = {'0': 'PRIMARY', '1': 'MED_POINT', '2': 'MED_AREA', '3': 'SURF_P', '4': 'SURF_L', '5': 'SURF_F', '6': 'GEN'}
for (i, k), m in per.items():
    mm = re.search(r'hbm_stage_kernel<(\d+), (\d+)>', k)
    name = names[mm.group(2)] if mm else k.split('(')[0][-24:]
    for a, b in m.items(): agg[name][a] += b
    cnt[name] += 1
tot = sum(v['gpu__time_duration.sum'] for v in agg.values())
print("%-16s %5s %10s %7s %9s %9s %8s %6s %10s" % ("kernel", "n", "avg us", "share", "dramR MB", "dramW MB", "issue %", "lanes", "DRAM GB/s"))
for n, v in sorted(agg.items(), key=lambda x: -x[1]['gpu__time_duration.sum']):
    c = cnt[n]; t = v['gpu__time_duration.sum']
    print("%-16s %5d %10.1f %6.1f%% %9.1f %9.1f %8.1f %6.1f %10.0f" % (n, c, t / c / 1e3, 100 * t / tot, v['dram__bytes_read.sum'] / c / 1e6, v['dram__bytes_write.sum'] / c / 1e6,
          v['smsp__issue_active.avg.pct_of_peak_sustained_active'] / c, v['smsp__thread_inst_executed_per_inst_executed.ratio'] / c, (v['dram__bytes_read.sum'] + v['dram__bytes_write.sum']) / t))
print("sum of kernel time per round: %.1f us; DRAM bytes per round: %.1f MB" % (tot / 1e3 / max(cnt['PRIMARY'], 1), sum(v['dram__bytes_read.sum'] + v['dram__bytes_write.sum'] for v in agg.values()) / 1e6 / max(cnt['PRIMARY'], 1)))
