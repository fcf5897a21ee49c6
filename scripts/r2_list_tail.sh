#!/bin/bash
# launch list (gpu__time_duration) of ~1.5 late k-steps only: skip the first launches instead of profiling them
SKIP=${1:-6000}; CNT=${2:-170}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s $SKIP -c $CNT --csv --log-file gpurun_out/tail.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --compress-tiles 0 --no-strong --no-cholesky > gpurun_out/tail.log 2>&1
echo "ncu rc=$?"
python scripts/dump_launches.py gpurun_out/tail.csv $CNT | cut -c1-150
