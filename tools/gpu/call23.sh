#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "distance or step or python" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "16 or distance or config2" 2>&1 | tail -3
timeout 400 python tools/dist_dbg.py 4096 shards 2>&1 | cut -c1-110 | tee gpurun_out/dist_dbg23.txt
timeout 200 python tools/dist_dbg.py 0 2>&1 | cut -c1-110 | tee gpurun_out/dist_dbg23_config2.txt
