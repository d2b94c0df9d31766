# round 2, first GPU pass: host facts, parity tests (all failures listed), short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
(nproc; free -g; df -h /dev/shm /tmp | tail -2) > gpurun_out/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/bench.err; cat gpurun_out/r02a_bench.json; cat gpurun_out/host.txt
