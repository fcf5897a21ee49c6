#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --compress-tiles 0 --no-strong --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], {k:round(v['ms_per_step'],1) for k,v in d['phases'].items()}, d.get('jacobi_or_bound_flags'))
PY
tail -3 gpurun_out/bench.err
bash scripts/ncu_step.sh r02_launches_b > /dev/null
N=$(grep -c gpu__time_duration gpurun_out/r02_launches_b.csv)
python scripts/dump_launches.py gpurun_out/r02_launches_b.csv $((N/48)) > gpurun_out/r02_launches_b_lastk_list.txt
grep -E "gemm" gpurun_out/r02_launches_b_lastk_list.txt
