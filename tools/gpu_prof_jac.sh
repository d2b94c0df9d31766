mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-jac_rx_kernel} -s ${SKIP:-1} -c 1 -o gpurun_out/${TAG:-prof_jac} -f python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_${TAG:-prof_jac}.log 2>&1
tail -2 gpurun_out/ncu_${TAG:-prof_jac}.log
