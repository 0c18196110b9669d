#!/bin/bash
# The multi-GPU measurements of round 2 on ONE box (run under `gpurun --gpus 8`): c5 strong scaling (global batch fixed),
# c2 weak scaling, the multi-rank correctness tests.  Results land in gpurun_out/.
mkdir -p gpurun_out
B=${C5_BATCH:-8192}
for N in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --config c5 --scaling strong --batch $B --gpus $N --steps 4 --warmup 3 > gpurun_out/r2_c5_strong_b${B}_n$N.json 2> gpurun_out/r2_c5_strong_b${B}_n$N.err
  tail -c 300 gpurun_out/r2_c5_strong_b${B}_n$N.err
done
# the largest global batch one GPU can hold (12288 users: 441 ms at N = 1), at N = 8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 \
  bench.py --config c5 --scaling strong --batch 12288 --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2_c5_strong_b12288_n8.json 2> gpurun_out/r2_c5_strong_b12288_n8.err
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_c2_weak_n$N.json 2> gpurun_out/r2_c2_weak_n$N.err
  tail -c 200 gpurun_out/r2_c2_weak_n$N.err
done
(timeout 500 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -rf 2>&1 | tail -30) > gpurun_out/r2_mg_tests_8gpu.log
tail -3 gpurun_out/r2_mg_tests_8gpu.log
