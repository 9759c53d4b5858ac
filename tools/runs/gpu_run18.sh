#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench_r1_n1.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench2.err; echo "bench2 rc=$?"; cut -c1-400 gpurun_out/bench_r1_n2.json; tail -3 gpurun_out/bench2.err
