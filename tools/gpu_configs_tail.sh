#!/bin/bash
# BASELINE cfg 4 as specified (exactly 10^4 steps) for the largest N with the shipped kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
timeout 1200 python tools/bench_configs.py --skip-cfg3 --nmin ${NMIN:-13} --nmax ${NMAX:-16} > gpurun_out/r02z_configs_n13_16.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
tail -n 3 gpurun_out/configs.err; cut -c1-500 gpurun_out/r02z_configs_n13_16.jsonl
