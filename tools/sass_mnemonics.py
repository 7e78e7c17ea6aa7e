#!/usr/bin/env python
"""Mnemonic histogram of the hot kernels in the built library (evidence of which pipes they use):
   python tools/sass_mnemonics.py [neuralgraphpde.jl_b200/libngpde.so] > profiles/rNN_sass_hist.txt
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA bulk copy), UTCBAR = tcgen05.commit,
HMMA = warp-level mma.sync (the persistent ODE kernels), ATOMS/UTCATOMSWS = shared-memory mbarrier / TMEM-allocator operations
(no global atomics: ATOMG / RED would show up here)."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neuralgraphpde.jl_b200", "libngpde.so")
KERNELS = ["mp_bwd_tc_kernelILb0", "mp_bwd_tc_kernelILb1", "mp_fwd_tc_kernelILb0", "mp_fwd_tc_kernelILb1", "gcn_aggregate_v4_kernelILi16ELb0",
           "gno_gemm_tc_kernel", "gno_gemm_tc_nloop_kernel", "edgeconv_ode_fwd_mma_kernelILi3", "edgeconv_ode_bwd_mma_kernelILi3",
           "hoist_fold_kernel", "dx_combine_kernel"]
KEEP = re.compile(r"^(FFMA|FMUL2|FADD2|HMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTCBAR|UTCATOMSWS|UBLKCP|UTMALDG|LDG|STG|LDS|STS|LDSM|MUFU|SHFL|SYNCS|ATOMS|ATOMG|ATOM|RED|REDUX|BAR|UCGABAR_ARV|UCGABAR_WAIT|MEMBAR|LDGSTS)$")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, hist = None, collections.defaultdict(collections.Counter)
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and fn:
        op = m.group(1)
        if KEEP.match(op):
            hist[fn][op] += 1
for k in KERNELS:
    for fn_name, h in hist.items():
        if k in fn_name:
            print("== " + k + ("  (" + fn_name[:70] + ")" if len(KERNELS) else ""))
            print("   " + "  ".join(f"{v:5d} {op}" for op, v in sorted(h.items(), key=lambda kv: -kv[1])))
