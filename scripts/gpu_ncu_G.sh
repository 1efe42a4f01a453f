# ncu --set full of one k_score launch on the fully assembled ~1 Gb workload; $1 = rigid pruning 0/1
R=${1:-0}
timeout 1500 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:k_score -s 6 -c 1 -f -o gpurun_out/r1_score_G_rigid$R python bench.py --workload G --start true --steps 8 --warmup 3 --flush-l2 0 --no-cpu-baseline --no-ref-gpu --rigid-pruning $R > gpurun_out/ncu_G.log 2>&1
tail -c 300 gpurun_out/ncu_G.log
