#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): launch list of one UNet+ControlNet step with DRAM bytes, `--set full` captures of the
# dominant and of the memory-bound kernels.  Outputs under gpurun_out/ (summaries are copied to profiles/ by hand).
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_mb32.csv \
    python tools/profile_step.py --mb 32 --steps 1 > gpurun_out/r2_launches_step_mb32.log 2>&1
for k in attn_tc_kernel xattn_tc_kernel gn_fused_kernel gn_reg_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/r2_full_$k \
      python tools/profile_step.py --mb 32 --steps 1 > gpurun_out/r2_full_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 30 -c 3 -f -o gpurun_out/r2_full_gemm_tc_kernel \
    python tools/profile_step.py --mb 32 --steps 1 > gpurun_out/r2_full_gemm_tc_kernel.log 2>&1
for k in canny_nms_kernel canny_hysteresis resample_pass_kernel layernorm_sub_kernel crop_normalize_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r2_full_$k \
      python tools/membound_once.py > gpurun_out/r2_full_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
