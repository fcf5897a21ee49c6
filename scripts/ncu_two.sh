#!/bin/bash
# ncu --set full of one late launch of two kernels (one bench run each). usage: bash scripts/ncu_two.sh K1 SKIP1 K2 SKIP2
bash scripts/ncu_kernel.sh "$1" "$2" "r01_$1" --warmup 0 --no-e2e
bash scripts/ncu_kernel.sh "$3" "$4" "r01_$3" --warmup 0 --no-e2e
