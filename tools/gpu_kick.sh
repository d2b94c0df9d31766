#!/bin/bash
# fast-kick pairs for N > 8: parity tests, then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale || exit 1
timeout 900 python -m pytest tests -q -m gpu -x -k "kick" --durations=5 > gpurun_out/pytest_kick.log 2>&1
echo "kick rc=$?"; tail -25 gpurun_out/pytest_kick.log
timeout 1200 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "all rc=$?"; tail -8 gpurun_out/pytest_gpu.log
