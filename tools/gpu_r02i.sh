# round 2, GPU pass i: production built without split-compile; cfg 3 regression hunt (A/B of builds); parity suite; bench
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
for lib in ab/libnbg_r01.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so ab/libnbg_u4.so ab/libnbg_pre_adjoint.so; do
  timeout 300 python tools/ab_cfg3.py $lib 20000 2>&1 | tail -n 1 | tee -a gpurun_out/r02i_ab_cfg3.jsonl
done
for lib in nbodygradient.jl_b200/csrc/libnbgrad_b200.so ab/libnbg_u4.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so; do
  timeout 200 python tools/ab_time.py $lib 4 2>&1 | tail -n 1 | tee -a gpurun_out/r02i_ab.jsonl
done
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
timeout 400 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 3 gpurun_out/bench.err; cat gpurun_out/r02i_bench.json
