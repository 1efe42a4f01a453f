timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score -s 3 -c 1 -f -o gpurun_out/r2_score_G_true_v2 python scripts/gpu_ncu_target.py --workload G --state true --steps 5 --nuis 0 > gpurun_out/ncu_score2.log 2>&1
tail -2 gpurun_out/ncu_score2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lnz_stream|k_lnz_refresh" -s 2 -c 2 -f -o gpurun_out/r2_lnz_G_true_v2 python scripts/gpu_ncu_target.py --workload G --state true --steps 3 --nuis 3 > gpurun_out/ncu_lnz2.log 2>&1
tail -2 gpurun_out/ncu_lnz2.log
