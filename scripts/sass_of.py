"""Print the SASS of the kernels whose mangled name contains every given substring, with an opcode histogram.

    python scripts/sass_of.py p2g_bulk3 [--hist] [--lib femflow_b200/_lib/libfemflow_mpm.so]
"""
import collections
import re
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
hist = "--hist" in sys.argv
lib = "femflow_b200/_lib/libfemflow_mpm.so"
for a in sys.argv[1:]:
    if a.startswith("--lib="):
        lib = a.split("=", 1)[1]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"(?=\t\tFunction : )", txt)
for b in blocks:
    m = re.match(r"\t\tFunction : (\S+)", b)
    if not m or not all(a in m.group(1) for a in args):
        continue
    ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", b, re.M)
    print(f"== {m.group(1)}: {len(ops)} instructions")
    if hist:
        for op, c in collections.Counter(ops).most_common(40):
            print(f"  {c:5d} {op}")
    else:
        print(b)
