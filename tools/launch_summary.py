"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel: time share per kernel and, when the DRAM metrics are present, bytes read / written per kernel and in total.
   usage: python tools/launch_summary.py launches.csv [rows]"""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0]); tot = 0.0; rd = wr = 0.0
seen = set()


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']; metric = row['Metric Name']
    name = row['Kernel Name']
    m = re.search(r'gemm_tc_kernel<\(bool\)(\d), \(bool\)(\d), \(int\)(\d+), \(int\)\d+, mcrn::(\w+)>', name)
    if m: key = f"tc<A_K={m.group(1)},B_K={m.group(2)},BN={m.group(3)},{m.group(4)}>"
    else:
        m = re.search(r'gemm_simt_kernel<mcrn::(\w+)>', name)
        key = f"simt<{m.group(1)}>" if m else re.sub(r'\(.*$', '', name)[-56:]
    if metric == 'gpu__time_duration.sum':
        v = v / 1000.0 if unit == 'ns' else (v * 1000.0 if unit == 'ms' else v)
        agg[key][1] += v; tot += v
        if row['ID'] not in seen:
            seen.add(row['ID']); agg[key][0] += 1
    elif metric == 'dram__bytes_read.sum':
        b = to_bytes(v, unit); agg[key][2] += b; rd += b
    elif metric == 'dram__bytes_write.sum':
        b = to_bytes(v, unit); agg[key][3] += b; wr += b
print(f"total {tot/1000:.2f} ms over {sum(a[0] for a in agg.values())} launches" + (f"; DRAM read {rd/1e6:.1f} MB, written {wr/1e6:.1f} MB" if rd + wr else ""))
for k, (n, t, r, w) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]:
    extra = f"  dram rd {r/n/1e6:7.2f} MB wr {w/n/1e6:7.2f} MB /launch" if rd + wr else ""
    print(f"{t/1000:8.3f} ms {100*t/tot:5.1f}%  n={n:4d}  avg {t/n:8.1f} us{extra}  {k}")
