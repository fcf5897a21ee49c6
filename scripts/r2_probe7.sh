#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --compress-tiles 0 --no-cpu-baseline --no-cholesky --strong-parity-tiles 16 "$@" > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], {k:round(v['ms_per_step'],1) for k,v in d['phases'].items()}, d.get('jacobi_or_bound_flags'))
s=d.get('strong_scaling',{}); print('strong', s.get('ms_per_step'), s.get('parity'), s.get('error'))
PY
tail -3 gpurun_out/bench.err
