#!/bin/bash
# One profiling pass on the GPU box (run under gpurun from the repo root):  tools/gpu_profile_pass.sh <tag>
# Produces under gpurun_out/: <tag>_pytest_gpu.log, <tag>_bench.json, <tag>_launches.csv (ncu launch list of the bench
# command), <tag>_c3_full.ncu-rep (--set full, one launch of each fused C3 kernel), <tag>_gcn_full.ncu-rep (GCN aggregate
# kernels inside the C5 chain, 64 graphs), <tag>_bench_c5.json, <tag>_bench_c2.json, <tag>_bench_c1.json
tag=${1:-r02}
out=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python tools/show_bench.py $out/${tag}_bench.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-c4 --no-cuda-graph > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mp_(fwd|bwd)_tc_kernel' -s 8 -c 4 -f -o $out/${tag}_c3_full \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong-c4 --no-cuda-graph > /dev/null 2> $out/${tag}_ncu_c3.err
timeout 400 ncu --set full --clock-control none -k regex:'gcn_aggregate' -s 12 -c 4 -f -o $out/${tag}_gcn_full \
    python bench.py --workload c5 --graphs 512 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> $out/${tag}_ncu_gcn.err
for wl in c5 c2 c1; do
  timeout 200 python bench.py --workload $wl --no-cpu-baseline > $out/${tag}_bench_${wl}.json 2>/dev/null
  python tools/show_bench.py $out/${tag}_bench_${wl}.json
done
timeout 200 python tools/ode_bench.py > $out/${tag}_ode_bench.log 2>&1
tail -4 $out/${tag}_ode_bench.log
timeout 200 python tools/tcb_phases_c5.py > $out/${tag}_phases_c5.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'edgeconv_ode' -c 2 -f -o $out/${tag}_ode_full \
    python tools/ode_bench.py > /dev/null 2> $out/${tag}_ncu_ode.err
timeout 400 ncu --set full --clock-control none -k regex:'mp_(fwd|bwd)_tc_kernel' -s 16 -c 8 -f -o $out/${tag}_c5_full \
    python bench.py --workload c5 --graphs 64 --steps 2 --warmup 2 --no-cpu-baseline > /dev/null 2> $out/${tag}_ncu_c5.err
sha256sum neuralgraphpde.jl_b200/libngpde.so > $out/${tag}_lib.sha256
ls -la $out/${tag}_*
