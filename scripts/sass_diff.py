"""Which kernels of a reference build of the library are byte-identical (SASS bodies, addresses and encodings
stripped) in the current build?  Used to show that adding opt-in kernel variants did not touch the code
generated for the default, measured kernels.

    python scripts/sass_diff.py <git-rev>      # builds csrc/ of that revision into /tmp and compares
"""
import collections
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    bodies, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            bodies[cur] = []
        elif cur and re.match(r"^\s+/\*[0-9a-f]{4}\*/", line):
            t = re.sub(r"/\* 0x[0-9a-f]* \*/", "", line)
            bodies[cur].append(re.sub(r"^\s*/\*[0-9a-f]*\*/", "", t).strip())
    return {k: (hashlib.md5("\n".join(v).encode()).hexdigest(), len(v)) for k, v in bodies.items()}


def main():
    rev = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "femflow_b200/csrc", "include"], capture_output=True, check=True)
        subprocess.run(["tar", "-x", "-C", tmp], input=tar.stdout, check=True)
        ref = os.path.join(tmp, "ref.so")
        subprocess.run(["nvcc", *FLAGS, "-o", ref, os.path.join(tmp, "femflow_b200/csrc/mpm_api.cu")], check=True,
                       capture_output=True)
        a = kernels(ref)
    b = kernels(os.path.join(ROOT, "femflow_b200", "_lib", "libfemflow_mpm.so"))
    have = collections.defaultdict(list)
    for k, (h, _) in b.items():
        have[h].append(k)
    changed = [(k, n) for k, (h, n) in a.items() if h not in have]
    print(f"{rev}: {len(a)} kernels, {len(a) - len(changed)} with an identical body in the current build "
          f"({len(b)} kernels)")
    for k, n in changed:
        print("  changed or gone:", k[:100], n, "instructions")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
