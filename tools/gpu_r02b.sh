# round 2, second GPU pass: all parity tests, the new bench line, the full-length runs
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 --durations=12 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|deviation|floor|capped|worst|rc=" gpurun_out/pytest_gpu.log | tail -40
timeout 400 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/r02b_bench.json
timeout 900 python bench.py --full --full-output chi2 > gpurun_out/r02b_bench_full_chi2.json 2> gpurun_out/bench_full_chi2.err; echo "full chi2 rc=$?"
tail -3 gpurun_out/bench_full_chi2.err; cat gpurun_out/r02b_bench_full_chi2.json
timeout 900 python bench.py --full --full-output arrays --nsys 16384 > gpurun_out/r02b_bench_full_arrays16k.json 2> gpurun_out/bench_full_arrays.err; echo "full arrays rc=$?"
tail -3 gpurun_out/bench_full_arrays.err; cat gpurun_out/r02b_bench_full_arrays16k.json
