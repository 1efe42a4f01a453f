# same-box A/B of the scoring kernel's knobs on the T workload: bash scripts/gpu_sweep_T.sh "IG_SPARSE_DIV=0" "IG_FORCE_SPLIT=24,4" ...
for cfg in "$@"; do
  env $cfg python bench.py --steps 2500 --warmup 300 --no-cpu-baseline --no-ref-gpu 2>&1 | tail -1 > gpurun_out/sweep.json
  echo -n "$cfg: "; python scripts/show_bench.py gpurun_out/sweep.json
done
