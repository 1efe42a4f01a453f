# A/B of an alternative build of the library (IG_B200_LIB) on T and on the assembled ~1 Gb workload
for lib in "$@"; do
  IG_B200_LIB=$lib python bench.py --steps 2500 --warmup 300 --no-cpu-baseline --no-ref-gpu 2>&1 | tail -1 > gpurun_out/ab.json; echo -n "$lib "; python scripts/show_bench.py gpurun_out/ab.json
  IG_B200_LIB=$lib python bench.py --workload G --start true --steps 200 --warmup 20 --flush-l2 0 --no-cpu-baseline --no-ref-gpu 2>&1 | tail -1 > gpurun_out/ab.json; echo -n "$lib "; python scripts/show_bench.py gpurun_out/ab.json
done
