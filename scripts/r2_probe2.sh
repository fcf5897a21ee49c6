#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --compress-tiles 0 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "rc=$?"; tail -c 2600 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
