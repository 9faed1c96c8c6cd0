#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "distance or mark or step or python" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
MMB_LIB=$PWD/dune-mmesh_b200/lib/dev/libmmesh_b200_dev.so timeout 400 python tools/dist_dbg.py 4096 2>&1 | tee gpurun_out/dist_dbg20.txt
MMB_LIB=$PWD/dune-mmesh_b200/lib/dev/libmmesh_b200_dev.so timeout 200 python tools/dist_dbg.py 0 2>&1 | tee gpurun_out/dist_dbg20_config2.txt
timeout 300 python bench.py --config 2 --steps 20 --no-cpu-baseline --no-e2e 2>/dev/null | tee gpurun_out/bench20_config2.json | cut -c1-200
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e 2>/dev/null | tee gpurun_out/bench20.json | cut -c1-200
