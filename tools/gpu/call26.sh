#!/bin/bash
# final single-GPU pass of the round: all GPU tests, the bench on configs 3 / 2 / 1, both arms, configs 4 / 5
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py 2> gpurun_out/bench26.err | tee gpurun_out/bench26.json | cut -c1-200; tail -2 gpurun_out/bench26.err
timeout 300 python bench.py --config 2 --steps 20 2>/dev/null | tee gpurun_out/bench26_config2.json | cut -c1-200
timeout 200 python bench.py --config 1 --steps 50 2>/dev/null | tee gpurun_out/bench26_config1.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/bench26_reference.json | cut -c1-200
timeout 200 python bench.py --impl reference --config 2 --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/bench26_reference_config2.json | cut -c1-200
timeout 100 python bench.py --impl reference --config 1 --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/bench26_reference_config1.json | cut -c1-200
timeout 300 python tools/kbench3d.py 110 5 2>&1 | tee gpurun_out/kbench3d_26.txt
timeout 300 python tools/kbench_config4.py 2>&1 | tee gpurun_out/kbench_config4_26.txt
