#!/bin/bash
# Round-2 closing run on ONE B200 (under gpurun): full GPU test suite, smoke, both bench arms, launch list with DRAM bytes and a
# `--set full` capture of the dominant kernel at this commit.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
T=${1:-final}
(timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6) > gpurun_out/r2_gputests_$T.txt 2>&1; cat gpurun_out/r2_gputests_$T.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/r2_smoke_$T.txt 2>&1; cat gpurun_out/r2_smoke_$T.txt
(timeout 900 python bench.py 2>gpurun_out/r2_bench_$T.err | tail -1) > gpurun_out/r2_bench_${T}_default.json; cat gpurun_out/r2_bench_${T}_default.json
(timeout 900 python bench.py --impl reference 2>gpurun_out/r2_bench_${T}_ref.err | tail -1) > gpurun_out/r2_bench_${T}_reference.json; cat gpurun_out/r2_bench_${T}_reference.json
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_${T}_step_mb32.csv \
    python tools/profile_step.py --mb 32 --steps 1 > gpurun_out/r2_launches_${T}_step_mb32.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_${T}_step_mb32.csv 30 --json gpurun_out/r2_launch_summary_${T}_step_mb32_dram.json > gpurun_out/r2_launches_${T}_step_mb32.txt 2>&1; head -12 gpurun_out/r2_launches_${T}_step_mb32.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 30 -c 3 -f -o gpurun_out/r2_full_${T}_gemm_tc_kernel \
    python tools/profile_step.py --mb 32 --steps 1 > gpurun_out/r2_full_${T}_gemm_tc_kernel.log 2>&1
ls -la gpurun_out/*${T}*.ncu-rep
