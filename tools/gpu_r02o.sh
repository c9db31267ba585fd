#!/bin/bash
# r02: k-NN kernel tests, timing at the MNIST shape, launch-level DRAM traffic of the tile kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn.py tests/test_gpu_recipes.py -m gpu -q --timeout=300 > gpurun_out/pytest_knn.log 2>&1
echo "pytest knn exit $?"; tail -5 gpurun_out/pytest_knn.log | cut -c1-400
timeout 600 python tools/knn_check.py full > gpurun_out/knn_full.log 2>&1
echo "knn full exit $?"; tail -1 gpurun_out/knn_full.log | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:knn_tile -s 3 -c 1 --csv --log-file gpurun_out/knn_tile_metrics.csv python tools/knn_check.py full > /dev/null 2>&1
tail -5 gpurun_out/knn_tile_metrics.csv | cut -c1-400
