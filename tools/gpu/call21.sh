#!/bin/bash
# 2 GPUs: the multi-GPU tests (incl. the C++ two-rank driver) and the N = 2 bench
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2> gpurun_out/bench21_n2.err | tee gpurun_out/bench21_n2.json | cut -c1-300
tail -3 gpurun_out/bench21_n2.err
