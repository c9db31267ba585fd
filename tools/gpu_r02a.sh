#!/bin/bash
# r02 first GPU pass: parity of the tile kernel (pytest -m gpu), A/B timings, sanitizer pass, ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python tools/kernel_ab.py c2 m134 --reps 12 > gpurun_out/ab_small.log 2>&1
echo "ab_small exit $?"; tail -20 gpurun_out/ab_small.log | cut -c1-600
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/kernel_ab.py c3 c5 > gpurun_out/ab_large.log 2>&1
echo "ab_large exit $?"; tail -8 gpurun_out/ab_large.log | cut -c1-600
timeout 600 python tools/kernel_ab.py c2 --reps 8 --variants "tiles:RB=12;tiles:RB=13;tiles:RB=11" > gpurun_out/ab_rb.log 2>&1
tail -4 gpurun_out/ab_rb.log | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "golden or zero or widths" --timeout=500 > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?"; tail -5 gpurun_out/sanitizer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:distortion_tile -s 3 -c 1 -o gpurun_out/r02_tile_c2 -f python tools/prof_target.py 8 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1500
