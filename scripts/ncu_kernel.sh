#!/bin/bash
# ncu --set full capture of ONE launch of a kernel inside the bench workload.
# usage: bash scripts/ncu_kernel.sh <kernel-regex> <skip> <outname> [bench args...]
K=$1; SKIP=$2; OUT=$3; shift 3
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 2 -f -o gpurun_out/$OUT \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-strong --no-cholesky "$@" > gpurun_out/${OUT}.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${OUT}.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py > gpurun_out/${OUT}_summary.txt
cat gpurun_out/${OUT}_summary.txt
