"""Group the SASS lines of one kernel (ncu --page source --csv) by execution count -- a cheap
way to separate loop nests -- and print each group's share of instructions, stall samples and
its stall-reason mix.  usage: ncu_stalls_by_region.py source.csv [min_share_pct]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
ia, ie, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
seen, uniq = set(), []
for r in body:
    if r[ia] not in seen:
        seen.add(r[ia])
        uniq.append(r)
tot = sum(int(r[ie]) for r in uniq)
tots = sum(int(r[iss]) for r in uniq)
groups = collections.defaultdict(list)
for r in uniq:
    groups[int(r[ie])].append(r)
floor = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
print(f"total warp instructions {tot}, samples {tots}")
allmix = collections.Counter()
for cnt, rs in sorted(groups.items(), key=lambda t: -sum(int(r[iss]) for r in t[1])):
    inst = sum(int(r[ie]) for r in rs)
    smp = sum(int(r[iss]) for r in rs)
    mix = collections.Counter()
    for r in rs:
        for h, i in stall:
            if r[i]:
                mix[h] += int(r[i])
    allmix.update(mix)
    if 100 * smp / tots < floor:
        continue
    top = ", ".join(f"{h[6:]} {100 * v / max(smp, 1):.0f}%" for h, v in mix.most_common(6))
    print(f"exec {cnt:>9} x {len(rs):>4} instrs: {100 * inst / tot:5.1f}% inst {100 * smp / tots:5.1f}% samples | {top}")
print("kernel:", ", ".join(f"{h[6:]} {100 * v / tots:.0f}%" for h, v in allmix.most_common(8)))
