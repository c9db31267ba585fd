#!/bin/bash
# r02: --set full capture of the late-epilogue head kernel (warm, a step that starts an iteration)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:step_head -s 30 -c 1 -o gpurun_out/r02_step_head -f python tools/prof_target.py 40 > gpurun_out/ncu_head.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_head.log
