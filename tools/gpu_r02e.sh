# round 2, GPU pass e: parity tests, A/B of builds, ncu captures (jac_rx full set + launch list), SASS listing
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|deviation from|capped|worst|RMS|rc=|trial" gpurun_out/pytest_gpu.log | cut -c1-500 | tail -30
for lib in ab/libnbg_pre_adjoint.so ab/libnbg_r02c_smem.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so ab/libnbg_r02c_smem.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so; do
  timeout 200 python tools/ab_time.py $lib 4 2>&1 | tail -n 1 | tee -a gpurun_out/r02e_ab.jsonl
done
timeout 400 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 3 gpurun_out/bench.err; cat gpurun_out/r02e_bench.json
R=r02e
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
LIB=nbodygradient.jl_b200/csrc/libnbgrad_b200.so
for k in jac_rx_kernel transit_adjoint_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o /tmp/${R}_$k python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
  python tools/ncu_summary.py /tmp/${R}_$k.ncu-rep > gpurun_out/${R}_$k.txt 2>&1
  python tools/ncu_hot.py /tmp/${R}_$k.ncu-rep $LIB $k 30 2>&1 | cut -c1-220 > gpurun_out/${R}_${k}_hot_lines.txt
done
python tools/ncu_profile_json.py /tmp/${R}_jac_rx_kernel.ncu-rep 32 gpurun_out/r02_jac_rx_profile.json
cat gpurun_out/${R}_jac_rx_kernel.txt
