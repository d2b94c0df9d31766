#!/bin/bash
# cfg 3 at the full 10^6 steps: per-segment device times, then one call
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale || exit 1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv -l 20 > gpurun_out/r02z_cfg3_smi.csv &
SMI=$!
timeout 1500 python tools/diag_cfg3.py > gpurun_out/r02z_cfg3_full.jsonl 2> gpurun_out/r02z_cfg3_full.err
echo "rc=$?"
kill $SMI
cat gpurun_out/r02z_cfg3_full.jsonl | cut -c1-300
tail -3 gpurun_out/r02z_cfg3_full.err
