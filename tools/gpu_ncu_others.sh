#!/bin/bash
# ncu --set full of the other kernels of the path with the final build (one launch each; 16,384 systems x 32-step window)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${ROUND:-r02z}
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
LIB=nbodygradient.jl_b200/csrc/libnbgrad_b200.so
for K in transit_kernel traj_kernel pair_op_kernel phi_dense_cached_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:^$K -s 1 -c 1 -f -o /tmp/${R}_$K python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$K.log 2>&1
  python tools/ncu_summary.py /tmp/${R}_$K.ncu-rep > gpurun_out/${R}_$K.txt 2>&1
  python tools/ncu_hot.py /tmp/${R}_$K.ncu-rep $LIB $K 20 2>&1 | cut -c1-220 > gpurun_out/${R}_${K}_hot_lines.txt
  head -30 gpurun_out/${R}_$K.txt | cut -c1-160
done
