# Round-end capture: tests, smoke, both bench arms, launch list (run under gpurun from the repo root)
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r1_pytest_gpu_final3.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r1_smoke_final3.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1_bench_reference_arm3.json
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r1_bench_final4.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_ncu_launches_final3.csv python bench.py --steps 20 --warmup 5 --no-push > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/r1_pytest_gpu_final3.log; tail -2 gpurun_out/r1_smoke_final3.log; cut -c1-300 gpurun_out/r1_bench_final4.json
