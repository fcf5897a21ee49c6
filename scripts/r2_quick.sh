#!/bin/bash
# GPU tests + a short headline bench (no CPU baseline / strong / cholesky legs): the edit-measure loop of round 2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong --no-cholesky --no-e2e --compress-tiles 0 "$@" > gpurun_out/bench_q.log 2>gpurun_out/bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_q.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms_per_step',d['ms_per_step'],'flags',d['jacobi_or_bound_flags'])
        print({k:round(v['ms_per_step'],1) for k,v in d['phases'].items()})
        e=d.get('effective',{})
        print('hist',e.get('jacobi_sweep_hist'),'lastk',e.get('jacobi_sweep_hist_last_k'), 'rank',e.get('c_rank_final_mean'))
        print('roofline',d.get('roofline'))
PY
tail -3 gpurun_out/bench_q.err
