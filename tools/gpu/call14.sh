#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_api.py -x -q -m gpu 2>&1 | tail -5
MMB_LIB=$PWD/dune-mmesh_b200/lib/dev/libmmesh_b200_dev.so timeout 300 python tools/proj_ab.py 4096 0,25,26,27,0 8 2>&1 | tee gpurun_out/proj_ab14.txt
