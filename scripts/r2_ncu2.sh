#!/bin/bash
# ncu --set full of the two block-Gram-Schmidt GEMMs (G = CU^T P, P -= CU G) of the last k-step of the first warm-up pass
# (17 k_gemm_dmma launches per k-step; the calibration pass comes first: 272 + 15 * 17 + 3)
bash scripts/ncu_kernel.sh k_gemm_dmma 530 r02_k_gemm_dmma_gs --no-e2e --compress-tiles 0 2>&1 | tail -80
