#!/bin/bash
mkdir -p gpurun_out
echo "== bench N=1 (full line)"; timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; echo "rc=$?"; tail -c 6000 gpurun_out/bench_n1.log; tail -5 gpurun_out/bench_n1.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_ref.log
