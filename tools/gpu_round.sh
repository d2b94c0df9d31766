# Full GPU round: parity tests, bench (both arms), ncu launch list, one ncu --set full capture per kernel.
set -x
mkdir -p gpurun_out
R=${ROUND:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
LIB=nbodygradient.jl_b200/csrc/libnbgrad_b200.so
for k in jac_rx_kernel traj_kernel transit_kernel pair_op_kernel phi_dense_kernel; do
  skip=1; [ $k = traj_kernel ] && skip=2; [ $k = phi_dense_kernel ] && skip=2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/${R}_$k python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
  python tools/ncu_summary.py /tmp/${R}_$k.ncu-rep > gpurun_out/${R}_$k.txt 2>&1
  python tools/ncu_hot.py /tmp/${R}_$k.ncu-rep $LIB $k 30 2>&1 | cut -c1-220 > gpurun_out/${R}_${k}_hot_lines.txt
done
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
