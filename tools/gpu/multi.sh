#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu/multi.sh N   -- multi-GPU tests (N >= 2) and the N-GPU bench
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-300
tail -3 gpurun_out/bench_n$N.err
