#!/bin/bash
# end-of-round evidence: launch list of one full headline step + ncu --set full of late Jacobi / rebuild-GEMM launches
bash scripts/r2_list.sh r02_launches_final | tail -25
bash scripts/ncu_kernel.sh k_jacobi_svd_rx 30 r02_final_k_jacobi_svd_rx --no-e2e --compress-tiles 0 2>&1 | grep -E "kernel =|gpu__time|dram__bytes|fp64_cycles|issue_active|warps_active|stalled_(barrier|wait|short|mio|long|math)" 
