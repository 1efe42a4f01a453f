# parity suite + short T bench (+ optional G benches with: bash scripts/gpu_check.sh G)
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 3000 --warmup 300 --no-cpu-baseline --no-ref-gpu 2>&1 | tail -1 > gpurun_out/bT.json
python scripts/show_bench.py gpurun_out/bT.json
if [ "$1" = "G" ]; then bash scripts/gpu_bench_G.sh; fi
