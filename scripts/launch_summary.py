"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: count, mean
duration and share of the summed kernel time, for the steady-state substeps (launches of the bench's
set-up phase -- torch fills / copies -- are dropped).  usage: launch_summary.py launches.csv [skip_first_n_p2g]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 1
names = [r[4] for r in rows]
# start after the `skip`-th P2G launch so that warm-up (first binning, full clears) is not mixed in
seen, start = 0, 0
for i, n in enumerate(names):
    if "p2g_" in n:
        seen += 1
        if seen == skip + 1:
            start = i
            break
acc = collections.defaultdict(list)
for r in rows[start:]:
    n = r[4]
    if n.startswith("void at::") or "at::native" in n:
        continue
    short = n.split("(")[0].replace("ffmpm::", "")
    acc[short].append(float(r[-1]) / 1e3)
tot = sum(sum(v) for v in acc.values())
for n, v in sorted(acc.items(), key=lambda t: -sum(t[1])):
    print(f"{n:<62} n={len(v):>3} avg {sum(v) / len(v):>9.1f} us  share {100 * sum(v) / tot:4.1f}%")
