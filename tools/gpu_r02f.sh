#!/bin/bash
# r02 scaling pass on an 8-GPU box: sharded bench at N = 8, 4, 2 (+ the 2-rank parity test).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/gpus8.txt 2>&1
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
  echo "bench N=$N exit $?"
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'timed_windows_ms')})
    print('single', d.get('single_gpu')); print('c2', d.get('c2_sharded')); print('parity', {k: v for k, v in d.get('parity', {}).items() if k != 'what'})
except Exception as e:
    print('no line', e); print(open('gpurun_out/bench_n$N.err').read()[-1500:])
PY
done
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout=500 > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi exit $?"; tail -4 gpurun_out/pytest_multi.log | cut -c1-300
