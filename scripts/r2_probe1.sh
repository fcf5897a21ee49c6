#!/bin/bash
# round 2, first GPU call: baseline state + accuracy-aware Jacobi stop (tests + bench) + BASELINE configs[3] shape on one GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
echo "== pytest default"; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_default.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_default.log
echo "== pytest acc stop"; HCB_JACOBI_ACC_STOP=1 timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_accstop.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_accstop.log
echo "== bench default"; timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --compress-tiles 0 > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_default.log
echo "== bench acc stop"; HCB_JACOBI_ACC_STOP=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --compress-tiles 0 > gpurun_out/bench_accstop.log 2>gpurun_out/bench_accstop.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_accstop.log
echo "== config 4 on one GPU"; HCB_JACOBI_ACC_STOP=1 timeout 900 python bench.py --tiles 32 --nb 2048 --acc 1e-6 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --compress-tiles 0 > gpurun_out/bench_cfg4_n1.log 2>gpurun_out/bench_cfg4_n1.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_cfg4_n1.log; tail -3 gpurun_out/bench_cfg4_n1.err
