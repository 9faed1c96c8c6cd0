#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 2> gpurun_out/bench25_n4.err | tee gpurun_out/bench25_n4.json | cut -c1-300
tail -2 gpurun_out/bench25_n4.err
