# ~1 Gb workload, fully assembled start (20 chromosomes): default (reference-faithful) and rigid-pruning modes
for r in 0 1; do
python bench.py --workload G --start true --steps 300 --warmup 20 --flush-l2 0 --no-cpu-baseline --no-ref-gpu --rigid-pruning $r 2>&1 | tail -1 > gpurun_out/bench_G_true_rigid$r.json
python -c "
import json; d=json.load(open('gpurun_out/bench_G_true_rigid$r.json')); print('rigid',$r,d['ms_per_step'],d['value'],d['roofline']['kernels']['k_score'], d['kernel_us_per_step'])"
done
