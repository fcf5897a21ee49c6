#!/bin/bash
# usage: bash scripts/r2_bench_n.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
echo "== bench N=$N"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 "$@" > gpurun_out/bench_n$N.log 2>gpurun_out/bench_n$N.err; echo "rc=$?"; tail -c 5000 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>gpurun_out/bench_ref_n$N.err; echo "rc=$?"; tail -c 1200 gpurun_out/bench_ref_n$N.log
