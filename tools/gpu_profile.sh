#!/bin/bash
# ncu launch lists (cold = default cache control, warm = --cache-control none) + optional full capture
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_cold.csv python tools/prof_target.py 12 ${PROF_CONS:-cen} > gpurun_out/prof_launch.log 2>&1
echo "cold list exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/launches_warm.csv python tools/prof_target.py 12 ${PROF_CONS:-cen} >> gpurun_out/prof_launch.log 2>&1
echo "warm list exit $?"
if [ -n "$PROF_FULL" ]; then
ncu --set full --clock-control none --import-source on -k regex:$PROF_FULL -s 4 -c 2 -o gpurun_out/prof_full -f python tools/prof_target.py 8 ${PROF_CONS:-cen} > gpurun_out/prof_full.log 2>&1
echo "full capture exit $?"
fi
