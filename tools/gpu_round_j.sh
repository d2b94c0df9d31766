# r01j: parity tests, e2e A/B of output slicing, cfg 4 sweep for N >= 9
mkdir -p gpurun_out
R=${ROUND:-r01j}
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_sliced.json 2> gpurun_out/bench_sliced.err; echo "bench rc=$?"
NBG_OUT_SLICES=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_whole.json 2> gpurun_out/bench_whole.err; echo "bench rc=$?"
NBG_OUT_SLICES=8 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_sliced8.json 2> gpurun_out/bench_sliced8.err; echo "bench rc=$?"
python - <<'PY'
import json
for t in ("sliced", "whole", "sliced8"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % t))
        print(t, "value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]))
    except Exception as ex:
        print(t, "failed", ex)
PY
timeout 300 python tools/bench_configs.py --skip-cfg3 --nmin 9 --generic > gpurun_out/${R}_configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
cut -c1-300 gpurun_out/${R}_configs.jsonl
tail -3 gpurun_out/bench_sliced.err
