#!/bin/bash
# end-of-round evidence: launch list (gpu__time_duration) of the headline command -- one warm-up pass + one timed pass
# (bench.py --profile: no calibration / trace passes, the rank bound of the calibrated run is given) -- and its summaries
OUT=r02_launches_final
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40000 --csv --log-file gpurun_out/$OUT.csv \
    python bench.py --profile --kc-bound 328 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --compress-tiles 0 --no-strong --no-cholesky > gpurun_out/$OUT.log 2>&1
echo "ncu rc=$?"
N=$(grep -c gpu__time_duration gpurun_out/$OUT.csv)
echo "launches: $N"
python scripts/summarize_launches.py gpurun_out/$OUT.csv $((N/2)) > gpurun_out/${OUT}_summary.txt       # the timed pass = second half
python scripts/dump_launches.py gpurun_out/$OUT.csv $((N/32)) > gpurun_out/${OUT}_lastk_list.txt          # its last k-step
head -30 gpurun_out/${OUT}_summary.txt
