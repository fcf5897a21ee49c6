#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest.log
bash scripts/ncu_step.sh r02_launches_a
N=$(grep -c gpu__time_duration gpurun_out/r02_launches_a.csv)
python scripts/dump_launches.py gpurun_out/r02_launches_a.csv $((N/48)) > gpurun_out/r02_launches_a_lastk_list.txt
