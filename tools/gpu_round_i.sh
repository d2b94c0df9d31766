# Round script (r01i): parity tests, cfg 3 / cfg 4 sweep, bench (both arms), smoke, ncu launch list. Ordered by priority; every
# command has its own timeout so that a slow one cannot eat the box.
mkdir -p gpurun_out
R=${ROUND:-r01i}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_configs.py > gpurun_out/${R}_configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
cut -c1-330 gpurun_out/${R}_configs.jsonl
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
