#!/bin/bash
# e2e (serial + pipelined) and the product-term preconditioner sweep cap
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-cholesky --compress-tiles 0 2>gpurun_out/p8.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default ms', d['ms_per_step'], 'e2e', json.dumps(d['e2e'])[:900])"
tail -2 gpurun_out/p8.err
for s in 3 5; do
HCB_PRECOND_SWEEPS=$s python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-cholesky --no-e2e --compress-tiles 0 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('precond sweeps $s ms', d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['phases'].items() if k in ('contraction','jacobi_svd')}, d['effective']['jacobi_sweep_hist'])"
done
