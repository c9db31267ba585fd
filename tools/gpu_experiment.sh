#!/bin/bash
# One gpurun call: GPU suite (default = mode 2), A/B timings, short bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout=300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_variants.py > gpurun_out/variants.log 2>&1; tail -12 gpurun_out/variants.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-100} --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
