# BASELINE.json configs[0] (yeast toy from the reference's tests/data): parity lockstep, GPU bench line, CPU arm (one cycle worth of steps)
timeout 600 python -m pytest tests -x -q -m gpu -k "yeast_toy" 2>&1 | tail -2
python bench.py --workload yeast_toy --steps 3000 --warmup 300 --no-ref-gpu > gpurun_out/bench_yeast_toy.json 2> gpurun_out/bench_yeast_toy.err; python scripts/show_bench.py gpurun_out/bench_yeast_toy.json
python bench.py --impl reference --workload yeast_toy --steps 64 --warmup 1 --cpu-budget-s 60 > gpurun_out/bench_yeast_toy_cpu_arm.json 2>> gpurun_out/bench_yeast_toy.err; cut -c1-300 gpurun_out/bench_yeast_toy_cpu_arm.json
