# Round-end capture (run under gpurun from the repo root): tests, smoke, both bench arms, launch lists, full ncu captures.
# Numbers printed by a run under ncu are never bench values; the bench lines come from the plain runs above them.
R=${1:-r2}
set -x
nvidia-smi -L > gpurun_out/${R}_gpu_box.txt; nvidia-smi --query-gpu=clocks.max.sm,clocks.sm,power.limit --format=csv >> gpurun_out/${R}_gpu_box.txt; nproc >> gpurun_out/${R}_gpu_box.txt
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 > gpurun_out/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${R}_bench_reference_arm.json
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/${R}_bench.json
python tools/bench_configs.py > gpurun_out/${R}_bench_configs.txt 2>&1
python tools/bench_gemm.py > gpurun_out/${R}_bench_gemm.txt 2>&1
python tools/time_backward.py > gpurun_out/${R}_time_backward.txt 2>&1
python tools/trace_gemm.py cfg2_image 1024 bf16 > gpurun_out/${R}_gemm_trace_cfg2.txt 2>&1
python tools/probe_chain_gemms.py 15 > gpurun_out/${R}_probe_chain_gemms.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${R}_ncu_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-push > gpurun_out/ncu_bench.log 2>&1
for c in "cfg2_image 1024 bf16" "cfg5_scaled 32 bf16" "cfg3_video_b1024 1024 fp32"; do set -- $c; ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_ncu_launches_$1_$3.csv python tools/run_cfg.py $1 $2 $3 2 > /dev/null 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_ncu_launches_backward_fp32.csv python tools/run_bwd.py 256 fp32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_tokens2 -s 3 -c 1 -o gpurun_out/${R}_k1_full python tools/run_one.py > gpurun_out/ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 5 -c 5 -o gpurun_out/${R}_tcgemm_cfg5 python tools/run_cfg.py cfg5_scaled 32 bf16 2 > gpurun_out/ncu_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 5 -c 5 -o gpurun_out/${R}_tcgemm_cfg2 python tools/run_cfg.py cfg2_image 1024 bf16 2 > gpurun_out/ncu_cfg2.log 2>&1
tail -2 gpurun_out/${R}_pytest_gpu.log; tail -2 gpurun_out/${R}_smoke.log; cut -c1-400 gpurun_out/${R}_bench.json
