"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1000.0 if unit == 'ns' else (v * 1000.0 if unit == 'ms' else v)
    name = row['Kernel Name']
    m = re.search(r'gemm_tc_kernel<\(bool\)(\d), \(bool\)(\d), \(int\)(\d+), \(int\)\d+, mcrn::(\w+)>', name)
    if m: key = f"tc<A_K={m.group(1)},B_K={m.group(2)},BN={m.group(3)},{m.group(4)}>"
    else:
        m = re.search(r'gemm_simt_kernel<mcrn::(\w+)>', name)
        key = f"simt<{m.group(1)}>" if m else name.split('(')[0][-48:]
    agg[key][0] += 1; agg[key][1] += v; tot += v
print(f"total {tot/1000:.2f} ms over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 22]:
    print(f"{t/1000:8.3f} ms {100*t/tot:5.1f}%  n={n:4d}  avg {t/n:8.1f} us  {k}")
