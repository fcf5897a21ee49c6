#!/bin/bash
# Launch list (gpu__time_duration per launch) of ONE full step of the bench workload + summaries of the whole step and
# of its last k iteration.  usage (under gpurun): bash scripts/ncu_step.sh [outname]
OUT=${1:-launches_step}
mkdir -p gpurun_out
timeout 2000 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40000 --csv --log-file gpurun_out/$OUT.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --compress-tiles 0 --no-strong --no-cholesky > gpurun_out/$OUT.log 2>&1
echo "ncu rc=$?"
N=$(grep -c gpu__time_duration gpurun_out/$OUT.csv)
echo "launches: $N"
# the run holds: calibration pass, the timed step, the per-k rank-trace pass -> the timed step is the middle third
python scripts/summarize_launches.py gpurun_out/$OUT.csv > gpurun_out/${OUT}_all_summary.txt
python scripts/summarize_launches.py gpurun_out/$OUT.csv $((N/3)) > gpurun_out/${OUT}_summary.txt
python scripts/summarize_launches.py gpurun_out/$OUT.csv $((N/48)) > gpurun_out/${OUT}_lastk_summary.txt
head -30 gpurun_out/${OUT}_summary.txt; head -30 gpurun_out/${OUT}_lastk_summary.txt
