#!/bin/bash
# final single-GPU pass of the round: all GPU tests, the bench on configs 3 / 2 / 1, both arms, configs 4 / 5
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py 2> gpurun_out/bench_pass.err | tee gpurun_out/bench_pass.json | cut -c1-200; tail -2 gpurun_out/bench_pass.err
timeout 300 python bench.py --config 2 --steps 20 2>/dev/null | tee gpurun_out/bench_pass_config2.json | cut -c1-200
timeout 200 python bench.py --config 1 --steps 50 2>/dev/null | tee gpurun_out/bench_pass_config1.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/bench_pass_reference.json | cut -c1-200
timeout 200 python bench.py --impl reference --config 2 --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/bench_pass_reference_config2.json | cut -c1-200
timeout 100 python bench.py --impl reference --config 1 --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/bench_pass_reference_config1.json | cut -c1-200
timeout 300 python tests/perf/kbench3d.py 110 5 2>&1 | tee gpurun_out/kbench3d_pass.txt
timeout 300 python tests/perf/kbench_config4.py 2>&1 | tee gpurun_out/kbench_config4_pass.txt
timeout 600 python tests/perf/flip_bench.py 4096 2>&1 | tee gpurun_out/flip_bench_4096.txt
