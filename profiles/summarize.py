"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel/grid -> markdown table."""
import collections, csv, re, sys
path = sys.argv[1]
lines = open(path).readlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
agg = collections.defaultdict(list)
for r in csv.DictReader(lines[start:]):
    if r['Metric Name'] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
    v = float(r['Metric Value'].replace(',', ''))
    v = v / 1e3 if r['Metric Unit'] == 'ns' else (v * 1e3 if r['Metric Unit'] == 'ms' else v)
    agg[(name, r['Grid Size'], r['Block Size'])].append(v)
T = sum(sum(v) for v in agg.values())
print("| share | avg us | launches | kernel | grid | block |\n|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print("| %.2f%% | %.2f | %d | `%s` | %s | %s |" % (100 * sum(v) / T, sum(v) / len(v), len(v), k[0], k[1], k[2]))
print("\ntotal %.1f us over %d launches" % (T, sum(len(v) for v in agg.values())))
