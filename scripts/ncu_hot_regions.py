"""Executed-instruction and stall-sample profile of one kernel from an `ncu --set full --import-source on` report,
grouped into contiguous SASS regions (split at the instructions whose text matches --split, e.g. barriers / loops).

    ncu -i prof.ncu-rep --page source --csv --print-source sass > src.csv
    python scripts/ncu_hot_regions.py src.csv g2p_tiled3 [--chunk 64]
"""
import csv
import sys

path, pat = sys.argv[1], sys.argv[2]
chunk = int(sys.argv[sys.argv.index("--chunk") + 1]) if "--chunk" in sys.argv else 64
rows = list(csv.reader(open(path)))
# the file holds one table per kernel: "Kernel Name" row, header row, then instruction rows
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        tables.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for t in tables:
    if pat not in t["name"]:
        continue
    h = t["hdr"]
    i_src, i_ex, i_th, i_smp = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    tot_ex = sum(int(r[i_ex]) for r in t["rows"])
    tot_smp = sum(int(r[i_smp]) for r in t["rows"])
    print(f"== {t['name'][:90]}\n   {len(t['rows'])} SASS instructions, {tot_ex} warp instructions executed, {tot_smp} stall samples")
    for a in range(0, len(t["rows"]), chunk):
        blk = t["rows"][a:a + chunk]
        ex = sum(int(r[i_ex]) for r in blk)
        smp = sum(int(r[i_smp]) for r in blk)
        ops = {}
        for r in blk:
            op = r[i_src].split()[0] if not r[i_src].strip().startswith("@") else r[i_src].split()[1]
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[i_ex])
        top = ", ".join(f"{k} {100 * v / max(ex, 1):.0f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
        print(f"   [{a:5d}..{a + len(blk):5d})  executed {100 * ex / tot_ex:5.1f} %  stall samples {100 * smp / max(tot_smp, 1):5.1f} %   {top}")
