#!/bin/bash
# launch list of one full headline step; prints the per-launch list of the last k-step
OUT=${1:-r02_launches_c}
bash scripts/ncu_step.sh $OUT > /dev/null
N=$(grep -c gpu__time_duration gpurun_out/$OUT.csv)
python scripts/dump_launches.py gpurun_out/$OUT.csv $((N/48)) > gpurun_out/${OUT}_lastk_list.txt
cat gpurun_out/${OUT}_lastk_summary.txt | head -30
