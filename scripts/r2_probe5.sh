#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest.log
echo "== bench (cholesky leg only extras)"; timeout 900 python bench.py --steps 2 --warmup 3 --no-e2e --compress-tiles 0 --no-strong --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'])
print(json.dumps(d.get('cholesky'), indent=1))
PY
tail -3 gpurun_out/bench.err
