set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r1_pytest_gpu_final2.log
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r1_bench_final2.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1_bench_reference_arm2.json
PASN_K1_PHASES=1 python tools/trace_k1.py > gpurun_out/r1_k1_phase_trace_serial.txt 2>&1
PASN_K1_PHASES=2 python tools/trace_k1.py > gpurun_out/r1_k1_phase_trace_two_phase.txt 2>&1
./tools/probe_gather.bin > gpurun_out/r1_probe_gather.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_ncu_launches_final2.csv python bench.py --steps 20 --warmup 5 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/r1_pytest_gpu_final2.log; cat gpurun_out/r1_bench_final2.json | cut -c1-400
