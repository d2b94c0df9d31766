set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jac_rx_kernel -s 1 -c 1 -o gpurun_out/r01_jac_rx python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_jac.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traj_kernel -s 1 -c 1 -o gpurun_out/r01_traj python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_traj.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:transit_kernel -s 1 -c 1 -o gpurun_out/r01_transit python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_transit.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
