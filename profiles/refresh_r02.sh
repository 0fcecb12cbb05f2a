set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02f_smoke.log 2>&1
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02f_ref.err | tail -1 > gpurun_out/r02f_bench_c3_reference.json
python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02f_bench.err | tail -1 > gpurun_out/r02f_bench_c3.json
python bench.py --config c4 --steps 100 --warmup 20 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02f_c4_full.json
python bench.py --config c4 --scale 0.125 --steps 100 --warmup 20 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02f_c4_shard512.json
cat gpurun_out/r02f_pytest.log gpurun_out/r02f_smoke.log; cut -c1-600 gpurun_out/r02f_bench_c3.json; cut -c1-300 gpurun_out/r02f_c4_full.json gpurun_out/r02f_c4_shard512.json
