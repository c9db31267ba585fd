#!/bin/bash
# One gpurun call: solver tests in all three modes, the whole GPU suite with mode 2 as the default, A/B timings.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -q --maxfail=20 --timeout=120 > gpurun_out/pytest_solver.log 2>&1
echo "pytest solver exit $?" >> gpurun_out/pytest_solver.log; tail -15 gpurun_out/pytest_solver.log
PYMDE_B200_SOLVER_MODE=2 timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest_mode2.log 2>&1
echo "pytest mode2 exit $?" >> gpurun_out/pytest_mode2.log; tail -6 gpurun_out/pytest_mode2.log
timeout 600 python tools/bench_variants.py > gpurun_out/variants.log 2>&1; tail -12 gpurun_out/variants.log
