# ncu --set full of the scoring kernels and of the cached-record likelihood at the MID-assembly state of workload G (the
# state of the headline bench line).  The burn-in (bomb + 2 cycles = 200 k steps) runs once OUTSIDE ncu and leaves the scaffold
# in /tmp; the captured processes start from it.
timeout 300 python scripts/gpu_ncu_target.py --workload G --state mid --save-state 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_stream|k_eval_flat|k_score" -s 6 -c 9 -f -o gpurun_out/r2_score_G_mid \
   python scripts/gpu_ncu_target.py --workload G --state file --steps 6 --nuis 0 > gpurun_out/ncu_score_mid.log 2>&1; tail -2 gpurun_out/ncu_score_mid.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_lnz_stream|k_lnz_refresh" -s 2 -c 2 -f -o gpurun_out/r2_lnz_G_mid \
   python scripts/gpu_ncu_target.py --workload G --state file --steps 3 --nuis 3 > gpurun_out/ncu_lnz_mid.log 2>&1; tail -2 gpurun_out/ncu_lnz_mid.log
ls -la gpurun_out/*.ncu-rep
