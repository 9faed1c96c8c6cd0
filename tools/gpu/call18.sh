#!/bin/bash
MMB_LIB=$PWD/dune-mmesh_b200/lib/dev/libmmesh_b200_dev.so timeout 400 python tools/dist_dbg.py 4096 shards 2>&1 | tee gpurun_out/dist_dbg18.txt
MMB_LIB=$PWD/dune-mmesh_b200/lib/dev/libmmesh_b200_dev.so timeout 200 python tools/dist_dbg.py 0 2>&1 | tee gpurun_out/dist_dbg18_config2.txt
