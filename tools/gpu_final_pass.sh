#!/bin/bash
# Lean end-of-round pass on the GPU box (run under gpurun from the repo root):  tools/gpu_final_pass.sh <tag>
# pytest -m gpu, the default bench line (+ reference arm), the C3 launch list and one --set full capture of each fused C3
# kernel (roofline.traffic), the other configs' bench lines, the library hash.
tag=${1:-r02}
out=gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python tools/show_bench.py $out/${tag}_bench.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-c4 --no-cuda-graph > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mp_(fwd|bwd)_tc_kernel' -s 8 -c 4 -f -o $out/${tag}_c3_full \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong-c4 --no-cuda-graph > /dev/null 2> $out/${tag}_ncu_c3.err
for wl in c2 c5 c1; do
  timeout 200 python bench.py --workload $wl --no-cpu-baseline > $out/${tag}_bench_${wl}.json 2>/dev/null
  python tools/show_bench.py $out/${tag}_bench_${wl}.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_c2_launches.csv \
    python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 400 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c4.json 2>/dev/null
python tools/show_bench.py $out/${tag}_bench_c4.json
sha256sum neuralgraphpde.jl_b200/libngpde.so > $out/${tag}_lib.sha256
ls -la $out/${tag}_*
