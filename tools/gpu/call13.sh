#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_api.py -x -q -m gpu -k "restore_delaunay or host_class_user_loop" 2>&1 | tail -15
timeout 300 python tools/flip_bench.py 1024 2>&1 | tee gpurun_out/flip_bench_1024.txt
timeout 600 python tools/flip_bench.py 4096 2>&1 | tee gpurun_out/flip_bench_4096.txt
timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench13.json | cut -c1-300
