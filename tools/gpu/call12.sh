#!/bin/bash
# full single-GPU measurement pass: gpu tests, bench (all configs), kernel A/B, reference arm
timeout 1300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py 2> gpurun_out/bench12.err | tee gpurun_out/bench12.json | cut -c1-200; tail -2 gpurun_out/bench12.err
MMB_LIB=$PWD/dune-mmesh_b200/lib/dev/libmmesh_b200_dev.so timeout 300 python tools/proj_ab.py 4096 0,24,25,0 8 2>&1 | tee gpurun_out/proj_ab12.txt
timeout 300 python bench.py --config 2 --steps 20 2>/dev/null | tee gpurun_out/bench12_config2.json | cut -c1-200
timeout 200 python bench.py --config 1 --steps 50 2>/dev/null | tee gpurun_out/bench12_config1.json | cut -c1-200
timeout 200 python tools/kbench3d.py 110 5 2>&1 | tee gpurun_out/kbench3d_12.txt
timeout 200 python tools/kbench_config4.py 2>&1 | tee gpurun_out/kbench_config4_12.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/bench12_reference.json | cut -c1-300
