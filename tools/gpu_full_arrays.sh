#!/bin/bash
# full-length run (26,667 steps) with every tt / dtdq0 / dtdelements row delivered to host arrays, final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
timeout 420 python bench.py --full --full-output arrays --nsys 16384 > gpurun_out/r02z_bench_full_arrays_16384.json 2> gpurun_out/bench_full_arrays.err; echo "full arrays rc=$?"
tail -n 2 gpurun_out/bench_full_arrays.err; cut -c1-1200 gpurun_out/r02z_bench_full_arrays_16384.json
