# round 2, GPU pass f: parity tests (slim scalar stream, plain transit dot product), A/B incl. experiment variants of the step loop
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|deviation from|capped|worst|RMS|rc=|trial" gpurun_out/pytest_gpu.log | cut -c1-500 | tail -30
for lib in ab/libnbg_r02c_smem.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so ab/libnbg_r02c_smem.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so; do
  timeout 200 python tools/ab_time.py $lib 4 2>&1 | tail -n 1 | tee -a gpurun_out/r02f_ab.jsonl
done
for v in 38 99 22 23 24 38 99; do
  NBG_RX_UNROLL=$v timeout 200 python tools/ab_time.py ab/libnbg_exp.so 4 2>&1 | tail -n 1 | tee -a gpurun_out/r02f_ab.jsonl
done
timeout 400 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 3 gpurun_out/bench.err; cat gpurun_out/r02f_bench.json
