# Round-end measurement batch (1 GPU): bench lines for every single-GPU config of BASELINE.json + ncu evidence.
# Outputs go to gpurun_out/final/ ; copy what is to be judged into profiles/.
O=gpurun_out/final; mkdir -p $O
python bench.py > $O/bench_T.json 2> $O/bench_T.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_T_reference_arm.json 2>> $O/bench_T.err
python bench.py --workload Y3 --steps 3000 --warmup 300 --no-cpu-baseline > $O/bench_Y3.json 2> $O/bench_Y3.err
python bench.py --workload yeast_toy --steps 3000 --warmup 300 --no-ref-gpu > $O/bench_yeast_toy.json 2>> $O/bench_Y3.err
python bench.py --impl reference --workload yeast_toy --steps 64 --warmup 1 --cpu-budget-s 60 > $O/bench_yeast_toy_cpu_arm.json 2>> $O/bench_Y3.err
python bench.py --workload G --start true --steps 300 --warmup 20 --flush-l2 0 --no-cpu-baseline --no-ref-gpu > $O/bench_G_true_start.json 2> $O/bench_G.err
python bench.py --workload G --start true --steps 300 --warmup 20 --flush-l2 0 --no-cpu-baseline --no-ref-gpu --rigid-pruning 1 > $O/bench_G_true_start_rigid.json 2>> $O/bench_G.err
python bench.py --workload G --burn-cycles 3 --steps 2000 --warmup 200 --no-cpu-baseline --ref-gpu-budget-s 20 > $O/bench_G_burn3.json 2>> $O/bench_G.err
# ncu: launch list at T (graph nodes), full sets of the scoring kernels at T (flat path: k_eval_flat, k_pick; benchmark state) and G (k_score, assembled)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 26000 -c 280 --csv --log-file $O/launches_T.csv python bench.py --steps 300 --warmup 50 --no-cpu-baseline --no-ref-gpu > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eval_flat -s 2100 -c 1 -f -o $O/ncu_k_eval_flat_T python bench.py --steps 200 --warmup 50 --graph 0 --no-cpu-baseline --no-ref-gpu > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_pick -s 2100 -c 1 -f -o $O/ncu_k_pick_T python bench.py --steps 200 --warmup 50 --graph 0 --no-cpu-baseline --no-ref-gpu > /dev/null 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:k_score -s 6 -c 1 -f -o $O/ncu_k_score_G python bench.py --workload G --start true --steps 8 --warmup 3 --flush-l2 0 --no-cpu-baseline --no-ref-gpu > /dev/null 2>&1
IG_B200_LIB=$PWD/instagraal_b200/libinstagraal_b200_tl.so python scripts/timeline.py T 2000 > $O/timeline_T.txt 2>&1
IG_B200_LIB=$PWD/instagraal_b200/libinstagraal_b200_tl.so python scripts/timeline.py Y3 2000 > $O/timeline_Y3.txt 2>&1
ls -la $O
