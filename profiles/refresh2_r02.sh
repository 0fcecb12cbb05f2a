#!/bin/bash
# Final refresh after the COLORED unit cap (CANNON_COLORED_UNIT_CONTACTS): remaining GPU tests, bench lines, phase trace,
# launch list and the --set full capture of the sweep. Run on the GPU box through gpurun.
set -x
python -m pytest tests/test_raycast.py tests/test_sph.py tests/test_trimesh.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02f_pytest_tail.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02f_ref.err | tail -1 > gpurun_out/r02f_bench_c3_reference.json
python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02f_bench.err | tail -1 > gpurun_out/r02f_bench_c3.json
python bench.py --config c4 --steps 100 --warmup 20 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02f_c4_full.json
python bench.py --config c4 --scale 0.125 --steps 100 --warmup 20 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02f_c4_shard512.json
python tools/gs_trace.py 20 c3 > gpurun_out/r02_gs_trace.log 2>&1
B="python bench.py --steps 12 --warmup 5 --no-e2e --no-cpu-baseline --no-c4"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none \
    -s 500 -c 230 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_gs_exact --launch-skip 9 --launch-count 1 -f -o gpurun_out/r02_k_gs_exact $B > gpurun_out/r02_ncu_k_gs_exact.log 2>&1
ncu -i gpurun_out/r02_k_gs_exact.ncu-rep --page raw --csv > gpurun_out/r02_k_gs_exact.raw.csv 2>/dev/null
cat gpurun_out/r02f_pytest_tail.log gpurun_out/r02f_smoke.log; cut -c1-400 gpurun_out/r02f_bench_c3.json
