#!/bin/bash
# Evidence capture for profiles/ (run on the GPU box through gpurun): headline bench line, ncu launch list of the
# contact-rich regime, one `--set full` capture each of the solver sweep and of the two launches of the pillar SAT kernel.
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 15000 -c 160 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gs_fast --launch-skip 215 --launch-count 1 -f -o gpurun_out/${TAG}_k_gs_fast \
    python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_gs.log 2>&1
# k_np_hull_warp<PILLAR, PHASE>: four launches per step (pillar sat, pillar clip, hull sat, hull clip); step 215
ncu --set full --clock-control none --import-source on -k regex:k_np_hull_warp --launch-skip 860 --launch-count 1 -f -o gpurun_out/${TAG}_k_np_hull_sat \
    python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_sat.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_np_hull_warp --launch-skip 861 --launch-count 1 -f -o gpurun_out/${TAG}_k_np_hull_clip \
    python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_clip.log 2>&1
ls -la gpurun_out/${TAG}_*
