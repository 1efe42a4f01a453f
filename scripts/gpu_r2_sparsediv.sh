for sd in ${SDS:-1 2 8 16}; do
IG_SPARSE_DIV=$sd python bench.py --workload G --start true --steps 400 --warmup 40 --chains 1 --nuisance-steps 10 --no-cpu-baseline --no-ref-gpu 2> gpurun_out/bench_G_sd.err | tail -1 > gpurun_out/bench_G_true_sd$sd.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_G_true_sd$sd.json'))
m=d['single_chain']['assembled']
print('sparse_div $sd: ms/step', round(m['ms_per_step'],4), 'scoring ms', round(m['kernels']['scoring']['ms_per_step'],4))
PY
done
