#!/bin/bash
# r02: late-epilogue solver steps: stage timing, solver tests, launch list of the step chain, bench
mkdir -p gpurun_out
timeout 300 python tools/solver_times.py 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_gpu_solver.py tests/test_gpu_state.py tests/test_gpu_configs.py -m gpu -q -x --timeout=900 > gpurun_out/pytest_solver.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_solver.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"step_|distortion_|project|tangent" -c 150 --csv --log-file gpurun_out/r02_launches_warm_late.csv python tools/prof_target.py 40 > gpurun_out/prof_launch.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_warm_late.csv
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_late1.json 2> gpurun_out/bench_late1.err
echo "bench1 exit $?"; cut -c1-300 gpurun_out/bench_late1.json; tail -3 gpurun_out/bench_late1.err
