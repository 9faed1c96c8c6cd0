#!/bin/bash
timeout 400 python tools/dist_dbg.py 4096 shards 2>&1 | cut -c1-120 | tee gpurun_out/dist_dbg24.txt
