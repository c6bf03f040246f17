#!/usr/bin/env python3
"""Per-source-line SASS instruction counts of one kernel in libaas_lmfb.so (needs -lineinfo).
usage: tools/sass_lines.py <kernel-substring> [top]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "aas_enhancement_b200", "libaas_lmfb.so")
pat = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], stdout=subprocess.PIPE,
                         text=True).stdout
cnt = collections.Counter()
fn, cur, total = None, "?", 0
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        fn = m.group(1)
        continue
    if line.strip().startswith(".section"):
        fn = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = os.path.basename(m.group(1)) + ":" + m.group(2)
        continue
    if fn and pat in fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line) and ".byte" not in line and ".dword" not in line:
        cnt[cur] += 1
        total += 1
print("kernel matching", pat, "instructions:", total)
for k, v in cnt.most_common(top):
    print(f"{v:6d}  {k}")
