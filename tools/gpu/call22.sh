#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 2> gpurun_out/bench22_n8.err | tee gpurun_out/bench22_n8.json | cut -c1-300
tail -2 gpurun_out/bench22_n8.err
