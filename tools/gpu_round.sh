# Final round script: parity tests, smoke, bench (both arms), ncu launch list.  ROUND names the outputs.
mkdir -p gpurun_out
R=${ROUND:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -10 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/bench_ref.err
NBG_OVERLAP=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/${R}_bench_serialized.json 2> gpurun_out/bench_ser.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/bench.err; cat gpurun_out/${R}_bench.json; cat gpurun_out/${R}_bench_reference.json
python - <<'PY'
import json, os
d = json.load(open("gpurun_out/%s_bench_serialized.json" % os.environ.get("ROUND", "r02a")))
print("serialized", "value %.4g" % d["value"], {k: round(v) for k, v in d["kernel_ms"].items()})
PY
