#!/bin/bash
# One gpurun call: full GPU tests (default build), solver tests under the chained-graph / PDL variants,
# then the A/B timings.  Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
MDE_B200_UNROLL=4 timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest_unroll4.log 2>&1
echo "pytest unroll4 exit $?" >> gpurun_out/pytest_unroll4.log; tail -3 gpurun_out/pytest_unroll4.log
MDE_B200_PDL=1 MDE_B200_UNROLL=4 timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest_pdl.log 2>&1
echo "pytest pdl exit $?" >> gpurun_out/pytest_pdl.log; tail -3 gpurun_out/pytest_pdl.log
timeout 600 python tools/bench_variants.py > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log | tail -12
timeout 600 python tools/bench_variants.py pdl > gpurun_out/variants_pdl.log 2>&1; cat gpurun_out/variants_pdl.log | tail -10
