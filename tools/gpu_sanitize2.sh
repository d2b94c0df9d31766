#!/bin/bash
# compute-sanitizer (memcheck, racecheck) on the paths added last: fast-kick Jacobian kernels for N > 8 (shared-memory plans up to 221 KB,
# first-kick scratch in shared / local memory) and the per-sample jac_step output
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
SEL="more_than_8_bodies or keeps_jac_step"
for tool in memcheck racecheck; do
  echo "=== $tool"
  timeout 700 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests -m gpu -q -k "$SEL" > gpurun_out/r02z_sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|hazard" gpurun_out/r02z_sanitize_$tool.log | sort | uniq -c | head -12
  grep -B2 -A12 "=========     at " gpurun_out/r02z_sanitize_$tool.log | head -40
done
