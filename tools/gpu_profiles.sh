# ncu launch list + one ncu --set full capture per kernel; text summaries only (reports are too big to bring back)
mkdir -p gpurun_out
R=${ROUND:-r01}
LIB=nbodygradient.jl_b200/csrc/libnbgrad_b200.so
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
for k in jac_rx_kernel traj_kernel transit_kernel pair_op_kernel phi_dense_kernel; do
  skip=1; [ $k = traj_kernel ] && skip=2; [ $k = phi_dense_kernel ] && skip=2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/${R}_$k python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
  python tools/ncu_summary.py /tmp/${R}_$k.ncu-rep > gpurun_out/${R}_$k.txt 2>&1
  python tools/ncu_hot.py /tmp/${R}_$k.ncu-rep $LIB $k 30 2>&1 | cut -c1-220 > gpurun_out/${R}_${k}_hot_lines.txt
  [ $k = jac_rx_kernel ] && cp /tmp/${R}_$k.ncu-rep gpurun_out/
done
ls -la gpurun_out | head -40
