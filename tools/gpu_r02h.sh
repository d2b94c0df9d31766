# round 2, GPU pass h: production = U4 pivot blocks; full parity suite, A/B, bench, the other BASELINE configurations
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|deviation from|capped|worst dev|RMS over|rc=|blocks" gpurun_out/pytest_gpu.log | cut -c1-400 | tail -30
for lib in nbodygradient.jl_b200/csrc/libnbgrad_b200.so ab/libnbg_u4.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so; do
  timeout 200 python tools/ab_time.py $lib 4 2>&1 | tail -n 1 | tee -a gpurun_out/r02h_ab.jsonl
done
timeout 400 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 3 gpurun_out/bench.err; cat gpurun_out/r02h_bench.json
timeout 1500 python tools/bench_configs.py --cfg3-steps 1000000 --nmin 2 --nmax 16 > gpurun_out/r02_configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
tail -n 3 gpurun_out/configs.err; cut -c1-330 gpurun_out/r02_configs.jsonl
