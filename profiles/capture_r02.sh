#!/bin/bash
# Round-2 evidence capture (run on the GPU box through gpurun): bench line, ncu launch list with DRAM bytes for EVERY
# kernel of the settled c3 step, `--set full` captures of the dominant kernels, launch list of a config-4 shard.
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 12 --warmup 5 --no-e2e --no-cpu-baseline --no-c4"
# per-launch time + DRAM traffic of every kernel: ~4 steps of the timed region (56 launches per step)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none \
    -s 500 -c 230 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launches.log 2>&1
for K in k_gs_exact k_np_hull_warp k_np_tasks k_rows_build k_schedule k_bp_small k_np_finalize k_units_build k_integrate k_prestep; do
  ncu --set full --clock-control none --import-source on -k regex:^$K --launch-skip 9 --launch-count 1 -f -o gpurun_out/${TAG}_$K $B > gpurun_out/${TAG}_ncu_$K.log 2>&1
  # gpurun brings back at most 64 MiB: keep the raw metric page of every capture, the report itself only for the sweeps
  ncu -i gpurun_out/${TAG}_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_$K.raw.csv 2>/dev/null
  [ $K = k_gs_exact ] || rm -f gpurun_out/${TAG}_$K.ncu-rep
done
# config 4, one 512-world shard
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 400 -c 110 --csv --log-file gpurun_out/${TAG}_launches_c4.csv \
    python bench.py --config c4 --scale 0.125 --steps 12 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_gs_world_exact --launch-skip 9 --launch-count 1 -f -o gpurun_out/${TAG}_k_gs_world_exact \
    python bench.py --config c4 --scale 0.125 --steps 12 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_k_gs_world_exact.log 2>&1
ncu -i gpurun_out/${TAG}_k_gs_world_exact.ncu-rep --page raw --csv > gpurun_out/${TAG}_k_gs_world_exact.raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_* | head -40
