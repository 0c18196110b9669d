#!/bin/bash
# The multi-GPU measurements of round 2 at N ranks of ONE box: `gpurun --gpus N -- bash tools/scale_run.sh N [prefix]`
# (one box per N: a call is charged N x its box time).  c5 strong scaling (global batch 8192 users), c2 weak scaling
# (4096 users per GPU), c4 row-sharded at N = 2; at N = 8 also the bulk-copy variant of the sharded gather and the
# peer-fabric micro-benchmark.  Results land in gpurun_out/.
N=${1:-8}; P=${2:-r2f}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$3" != "weak-only" ]; then
timeout 400 $TR --master-port 29601 bench.py --config c5 --scaling strong --batch 8192 --gpus $N --steps 4 --warmup 3 \
  > gpurun_out/${P}_c5_b8192_n$N.json 2> gpurun_out/${P}_c5_b8192_n$N.err
tail -c 200 gpurun_out/${P}_c5_b8192_n$N.err
fi
timeout 300 $TR --master-port 29602 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/${P}_c2_weak_n$N.json 2> gpurun_out/${P}_c2_weak_n$N.err
tail -c 200 gpurun_out/${P}_c2_weak_n$N.err
[ "$3" = "weak-only" ] && exit 0
if [ "$N" = "2" ]; then
  timeout 300 $TR --master-port 29603 bench.py --config c4 --gpus 2 --steps 10 --warmup 3 \
    > gpurun_out/${P}_c4_n2.json 2> gpurun_out/${P}_c4_n2.err
fi
if [ "$N" = "8" ]; then
  POI_PEER_GATHER=1 timeout 400 $TR --master-port 29604 bench.py --config c5 --scaling strong --batch 8192 --gpus $N --steps 4 --warmup 3 \
    > gpurun_out/${P}_c5_b8192_n${N}_bulk.json 2> gpurun_out/${P}_c5_b8192_n${N}_bulk.err
  timeout 200 $TR --master-port 29605 tools/peer_gather_bench.py 2>/dev/null | grep "^{" > gpurun_out/${P}_peer_n${N}_mode0.json
  cat gpurun_out/${P}_peer_n${N}_mode0.json
fi
ls -la gpurun_out/${P}_*n$N* | awk '{print $5, $9}'
