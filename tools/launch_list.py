#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: tools/launch_list.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rN_launches.txt"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.Counter()
cnt = collections.Counter()
for r in rows:
    tot[r[4]] += float(r[14].replace(",", ""))
    cnt[r[4]] += 1
s = sum(tot.values()) or 1.0
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("(per-launch times are serialised and cold-cache; the kernel's share of the step is what matters)")
for k, v in tot.most_common():
    print(f"{cnt[k]:4d} launches  {v / 1e3:12.1f} us total  {v / cnt[k] / 1e3:10.1f} us avg  {100 * v / s:5.1f}%  {k[:110]}")
