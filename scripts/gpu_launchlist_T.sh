# launch list of the T workload in graph mode (cold-cache, serialised: compare shares)
timeout 1000 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 26000 -c 280 --csv --log-file gpurun_out/launches_T.csv python bench.py --steps 300 --warmup 50 --no-cpu-baseline --no-ref-gpu > gpurun_out/l.log 2>&1
tail -c 200 gpurun_out/l.log
