"""Aggregate an ncu launch list (gpu__time_duration.sum CSV) by kernel: python agg_launches.py <csv> [skip_first_n_calls]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
cnt, tot = collections.Counter(), collections.Counter()
for r in rows[1:]:
    short = r[ki].split("(")[0].split("::")[-1]
    cnt[short] += 1
    tot[short] += float(r[vi].replace(",", ""))
for s in cnt:
    print(f"{s:24s} n={cnt[s]:5d} total={tot[s] / 1e3:10.1f} us  avg={tot[s] / cnt[s] / 1e3:9.2f} us")
