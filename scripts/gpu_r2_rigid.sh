# rigid-pruning mode at G assembled: bench + ncu of k_stream / k_eval_flat<list>; mid-state captures for the traffic table
python bench.py --workload G --start true --mode rigid --steps 600 --warmup 50 --chains 1 --no-cpu-baseline --no-ref-gpu 2> gpurun_out/bench_G_rigid.err | tail -1 > gpurun_out/bench_G_true_rigid_stream.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_G_true_rigid_stream.json'))
m=d['single_chain']['assembled']
print('rigid assembled: ms/step', round(m['ms_per_step'],4), 'value', round(m['value']), 'scoring', m['kernels']['scoring'], {k:round(v,1) for k,v in m['kernel_us_per_step'].items()})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stream|k_eval_flat" -s 8 -c 2 -f -o gpurun_out/r2_stream_G_true_rigid python scripts/gpu_ncu_target.py --workload G --state true --steps 6 --nuis 0 --rigid 1 > gpurun_out/ncu_stream_rigid.log 2>&1; tail -2 gpurun_out/ncu_stream_rigid.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_stream|k_eval_flat|k_score|k_lnz_stream|k_lnz_refresh" -s 25 -c 5 -f -o gpurun_out/r2_mid_G python scripts/gpu_ncu_target.py --workload G --state mid --steps 7 --nuis 3 > gpurun_out/ncu_mid.log 2>&1; tail -2 gpurun_out/ncu_mid.log
