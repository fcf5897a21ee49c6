#!/bin/bash
# usage: bash scripts/r2_tests.sh [pytest -k expression]
mkdir -p gpurun_out
if [ -n "$1" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "$1" > gpurun_out/pytest.log 2>&1
else
  timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1
fi
echo "rc=$?"; tail -40 gpurun_out/pytest.log
