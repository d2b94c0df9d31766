# round 2, third GPU pass: parity tests after the adjoint / difference-form fix, bench, A/B of the Jacobian kernel before / after the adjoint
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|deviation|floor|capped|worst|RMS|rc=" gpurun_out/pytest_gpu.log | tail -40
timeout 400 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 3 gpurun_out/bench.err; cat gpurun_out/r02c_bench.json
for lib in ab/libnbg_pre_adjoint.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so ab/libnbg_pre_adjoint.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so; do
  timeout 200 python tools/ab_time.py $lib 4 2>&1 | tail -n 1 | tee -a gpurun_out/r02c_ab.jsonl
done
