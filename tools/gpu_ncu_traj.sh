#!/bin/bash
# ncu --set full of the light trajectory kernel (traj_kernel<false,2>: the main-loop steps) with the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${ROUND:-r02z}
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
LIB=nbodygradient.jl_b200/csrc/libnbgrad_b200.so
timeout 150 ncu --set full --clock-control none --import-source on -k regex:^traj_kernel -s 2 -c 1 -f -o /tmp/${R}_traj_light python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_traj_light.log 2>&1
python tools/ncu_summary.py /tmp/${R}_traj_light.ncu-rep > gpurun_out/${R}_traj_kernel_light.txt 2>&1
python tools/ncu_hot.py /tmp/${R}_traj_light.ncu-rep $LIB traj_kernel 20 2>&1 | cut -c1-220 > gpurun_out/${R}_traj_kernel_light_hot_lines.txt
head -30 gpurun_out/${R}_traj_kernel_light.txt | cut -c1-160
