#!/bin/bash
# ncu --set full of one late launch of the rebuild GEMMs and of the Jacobi kernel inside the headline workload
bash scripts/ncu_kernel.sh k_gemm_dmma 797 r02_k_gemm_dmma_rebuild --no-e2e --compress-tiles 0 2>&1 | tail -45
bash scripts/ncu_kernel.sh k_jacobi_svd_rx 30 r02_k_jacobi_svd_rx --no-e2e --compress-tiles 0 2>&1 | tail -45
ls -la gpurun_out/*.ncu-rep | tail -3
