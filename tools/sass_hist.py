#!/usr/bin/env python
"""SASS instruction count per source line of one kernel (code-size / instruction-cache diagnosis).
   python tools/sass_hist.py neuralgraphpde.jl_b200/build/ngpde_tc.o mp_bwd_tc_kernelILb0 [top]"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
    out = ""
    for cubin in glob.glob(os.path.join(d, "*.cubin")):
        out += subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line, fn = None, None
hist, tot = collections.Counter(), collections.Counter()
for ln in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln) and fn:
        tot[fn[:60]] += 1
        if pat in fn:
            hist[line] += 1
for k, v in tot.items():
    print(f"{v:7d} instructions ({v * 16 // 1024} KB)  {k}")
for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:top]:
    print(v, k)
