#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_api.py -x -q -m gpu -k "restore_delaunay or host_class_user_loop" 2>&1 | tail -5
MMB_LIB=$PWD/dune-mmesh_b200/lib/dev/libmmesh_b200_dev.so timeout 300 python tools/proj_ab.py 4096 0,28,29,0 8 2>&1 | tee gpurun_out/proj_ab15.txt
