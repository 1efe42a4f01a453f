# final bench lines of the small workloads (T with both baselines, Y3 with the reference-GPU baseline, yeast toy)
O=gpurun_out/final; mkdir -p $O
python bench.py > $O/bench_T.json 2> $O/bench_T.err; python scripts/show_bench.py $O/bench_T.json
python bench.py --workload Y3 --steps 3000 --warmup 300 --no-cpu-baseline > $O/bench_Y3.json 2> $O/bench_Y3.err; python scripts/show_bench.py $O/bench_Y3.json
python bench.py --workload yeast_toy --steps 3000 --warmup 300 --no-ref-gpu > $O/bench_yeast_toy.json 2>> $O/bench_Y3.err; python scripts/show_bench.py $O/bench_yeast_toy.json
