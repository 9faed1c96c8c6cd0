#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "distance or mark or patch or step" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "distance or 16 or full" 2>&1 | tail -4
MMB_LIB=$PWD/dune-mmesh_b200/lib/dev/libmmesh_b200_dev.so timeout 300 python tools/dist_dbg.py 4096 2>&1 | tee gpurun_out/dist_dbg16.txt
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e 2>/dev/null | tee gpurun_out/bench16.json | cut -c1-200
