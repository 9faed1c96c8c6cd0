#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_python_api.py -x -q -m gpu -k "3d or three or tet or skeleton or trace" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "config5 or 3d" 2>&1 | tail -4
timeout 400 python tools/kbench3d.py 110 5 2>&1 | tee gpurun_out/kbench3d_27.txt
