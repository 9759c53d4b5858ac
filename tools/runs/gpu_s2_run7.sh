#!/bin/bash
# session-2 call 7 (2 GPUs): the N=2 bench arm exactly as the driver launches it (torchrun, one rank per GPU, DDP over NCCL)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_s2_n2.json 2> gpurun_out/bench_s2_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/bench_s2_n2.json | cut -c1-1200; tail -5 gpurun_out/bench_s2_n2.err
