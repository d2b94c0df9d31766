# r01k: e2e stage trace, chunk-length A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "one_shot or device_ic or trappist8_batch or fused" > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
NBG_TRACE=1 NBG_OUT_SLICES=8 timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; echo "bench rc=$?"
grep "nbg trace" gpurun_out/bench_trace.err | tail -8
NBG_TRACE=1 NBG_OUT_SLICES=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_trace1.json 2> gpurun_out/bench_trace1.err; echo "bench rc=$?"
grep "nbg trace" gpurun_out/bench_trace1.err | tail -4
for g in 70 100; do
NBG_OUT_SLICES=8 timeout 300 python bench.py --no-cpu-baseline --steps 3 --stream-budget-gb $g > gpurun_out/bench_budget$g.json 2> gpurun_out/bench_budget$g.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for t in ("trace", "trace1", "budget70", "budget100"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % t))
        print(t, "value %.4g e2e %.4g chunk %d" % (d["value"], d["e2e"]["value"], d["config"]["chunk_steps"]), {k: round(v) for k, v in d["kernel_ms"].items()})
    except Exception as ex:
        print(t, "failed", ex)
PY
