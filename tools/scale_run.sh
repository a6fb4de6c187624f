#!/bin/bash
# Strong-scaling run of the metric on one 8-GPU box: bench.py at N = 1, 2, 4, 8 (in-kernel peer-memory all-reduce), N = 8
# with the separate NCCL all-reduce, and the multi-GPU tests.  Outputs under gpurun_out/ (one JSON line per run).
#   gpurun --gpus 8 -- 'bash tools/scale_run.sh'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -q -m gpu 2>&1 | tail -2 > gpurun_out/scale_multigpu_tests.log
python bench.py --gpus 1 --steps 100 --warmup 10 --no-extra 2> gpurun_out/scale_n1.err | grep '^{' > gpurun_out/scale_n1.json
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 100 --warmup 10 --no-extra 2> gpurun_out/scale_n$n.err | grep '^{' > gpurun_out/scale_n$n.json
  echo "n=$n rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 8 --steps 100 --warmup 10 --no-extra --collective nccl 2> gpurun_out/scale_n8_nccl.err | grep '^{' > gpurun_out/scale_n8_nccl.json
cat gpurun_out/scale_multigpu_tests.log
