#!/bin/bash
# Round-2 result table: one bench line per BASELINE configuration (run on the GPU box through gpurun).
OUT=gpurun_out/${1:-r02_table}.jsonl
: > $OUT
python bench.py --config c1 --solver reference --steps 600 --warmup 20 --no-e2e | tail -1 >> $OUT
python bench.py --config c1 --solver colored --steps 600 --warmup 20 --no-e2e --no-cpu-baseline | tail -1 >> $OUT
python bench.py --config c2 --solver reference --steps 300 --warmup 20 --no-e2e | tail -1 >> $OUT
python bench.py --config c2 --solver colored --steps 300 --warmup 20 --no-e2e --no-cpu-baseline | tail -1 >> $OUT
python bench.py --config c3 --steps 100 --warmup 10 --no-c4 --no-cpu-baseline | tail -1 >> $OUT
python bench.py --config c3 --solver colored_f32 --steps 100 --warmup 10 --no-c4 --no-cpu-baseline --no-e2e | tail -1 >> $OUT
python bench.py --config c4 --solver reference --steps 300 --warmup 20 --no-e2e --no-cpu-baseline | tail -1 >> $OUT
python bench.py --config c4 --solver colored --steps 300 --warmup 20 --no-e2e --no-cpu-baseline | tail -1 >> $OUT
python bench.py --config c5 --steps 100 --warmup 200 --no-e2e --no-cpu-baseline | tail -1 >> $OUT
wc -l $OUT
