#!/bin/bash
# r02 multi-GPU pass (gpurun --gpus 2 | 8): sharded evaluation / solve parity, then the sharded bench.
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_state.py -m gpu -q -x --timeout=800 > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi exit $?"; tail -25 gpurun_out/pytest_multi.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-3500; tail -5 gpurun_out/bench_n$N.err | cut -c1-300
