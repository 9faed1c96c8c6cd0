#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_api.py -x -q -m gpu -k "restore_delaunay or host_class_user_loop" 2>&1 | tail -3
timeout 200 python tests/perf/flip_bench.py 1024 2>&1 | tee gpurun_out/flip_bench_1024.txt
timeout 300 python tests/perf/flip_bench.py 4096 0.3 nooracle 2>&1 | tee gpurun_out/flip_bench_4096.txt
