#!/bin/bash
# One gpurun call: smoke + GPU parity tests + a short bench + the ncu launch list. Everything lands in gpurun_out/.
# usage (under gpurun): bash scripts/gpu_check.sh [quick|full]
mkdir -p gpurun_out
MODE=${1:-full}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest.log
echo "== fp64 yardstick"; timeout 300 python scripts/fp64_peak.py > gpurun_out/fp64_peak.json 2>gpurun_out/fp64_peak.err; cat gpurun_out/fp64_peak.json
echo "== bench"; timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
if [ "$MODE" = "full" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 800 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --tiles 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc=$?"; python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt | head -30
fi
