#!/bin/bash
# r02: Standardized for 32 < m <= 256 on the device (tiled Gram + Newton-Schulz): kernel and solver tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver.py -m gpu -q --timeout=600 -k "wide or standardized or proj" > gpurun_out/pytest_wide.log 2>&1
echo "pytest wide exit $?"; tail -40 gpurun_out/pytest_wide.log | cut -c1-300
