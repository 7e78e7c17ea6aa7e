#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the C3 fused kernels (tools/gpu_profile_pass.sh):
DRAM bytes (read + write) per launch of each kernel, stamped with the sha256 of the libngpde.so that was profiled -- bench.py
reports `roofline.traffic` only when the benched build has the same hash.
    python tools/make_traffic.py gpurun_out/<tag>_c3_full.ncu-rep gpurun_out/<tag>_lib.sha256 [--gcn gpurun_out/<tag>_gcn.ncu-rep]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, data = r[0], r[1], r[2:]
    return hdr, units, data


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    rep, shafile = sys.argv[1], sys.argv[2]
    sha = open(shafile).read().split()[0]
    hdr, units, data = rows(rep)
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    names = {"mp_fwd_tc_kernel<0>": "mp_fwd_tc_kernel<edge>", "mp_fwd_tc_kernel<1>": "mp_fwd_tc_kernel<node>",
             "mp_bwd_tc_kernel<1,": "mp_bwd_tc_kernel<node>", "mp_bwd_tc_kernel<0,": "mp_bwd_tc_kernel<edge>"}
    c3 = {}
    for r in data:
        kn = r[ki].replace("(bool)", "").replace("void ", "").replace("ngpde::", "").replace(" ", "")
        for pat, nice in names.items():
            if kn.startswith(pat.replace(" ", "")):
                c3[nice] = to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
    out = {"_source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, one launch of each fused kernel, C3 VMHConv "
                      "256x256; dram__bytes_read.sum + dram__bytes_write.sum per launch)",
           "lib_sha256": sha, "c3": c3}
    if "--gcn" in sys.argv:
        grep = sys.argv[sys.argv.index("--gcn") + 1]
        hdr, units, data = rows(grep)
        ki, ri, wi, di = (hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"),
                          hdr.index("gpu__time_duration.sum"))
        g = []
        for r in data:
            b = to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
            us = float(r[di].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[di], 1.0)
            g.append({"kernel": r[ki].split("(")[0][-40:], "dram_bytes": b, "duration_us": us, "dram_gbs": b / us / 1e3})
        out["c5_gcn_aggregate"] = g
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
