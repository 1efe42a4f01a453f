# ncu --set full captures at G (assembled state): k_full_lnz (nuisance likelihood) and k_score (exact mode)
set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_full_lnz -s 2 -c 1 -f -o gpurun_out/r2_full_lnz_G_true python scripts/gpu_ncu_target.py --workload G --state true --steps 2 --nuis 4 > gpurun_out/ncu_full_lnz.log 2>&1
tail -3 gpurun_out/ncu_full_lnz.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score -s 3 -c 1 -f -o gpurun_out/r2_score_G_true python scripts/gpu_ncu_target.py --workload G --state true --steps 5 --nuis 0 > gpurun_out/ncu_score.log 2>&1
tail -3 gpurun_out/ncu_score.log
ls -la gpurun_out/*.ncu-rep
