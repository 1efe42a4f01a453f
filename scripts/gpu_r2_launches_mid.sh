# ncu launch list (gpu__time_duration per launch) of single-chain steps at the MID-assembly state of workload G; burn-in outside ncu
timeout 200 python scripts/gpu_ncu_target.py --workload G --state mid --save-state 2>&1 | tail -1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 520 --csv --log-file gpurun_out/r2_launches_G_mid.csv \
   python scripts/gpu_ncu_target.py --workload G --state file --steps 36 --nuis 0 > gpurun_out/l_mid.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/r2_launches_G_mid.csv | tail -22
