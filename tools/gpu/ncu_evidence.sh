#!/bin/bash
# evidence pass: plain run first, then the launch list and one --set full capture of the two dominant kernels
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/bench17_plain.json 2> gpurun_out/bench17_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launch_list_16M.csv $CMD > gpurun_out/ncu17a.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_project|k_distance2d" -c 4 -o gpurun_out/r02_dominant_16M $CMD > gpurun_out/ncu17b.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/ | tail -8
