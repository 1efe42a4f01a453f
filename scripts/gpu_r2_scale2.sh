# 2-GPU check of the replica path: NCCL tests + bench at N=2 (as the driver launches it)
timeout 600 python -m pytest tests/test_gpu_replicas.py -m gpu -q > gpurun_out/r2_tests_nccl.log 2>&1; tail -3 gpurun_out/r2_tests_nccl.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 100 > gpurun_out/bench_G_2gpu.json 2> gpurun_out/bench_G_2gpu.err; echo "rc=$?"; tail -c 400 gpurun_out/bench_G_2gpu.err; head -c 900 gpurun_out/bench_G_2gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_G_refarm.json 2> gpurun_out/bench_G_refarm.err; echo "rc=$?"; head -c 600 gpurun_out/bench_G_refarm.json
