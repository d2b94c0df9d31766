# r01m: DMMA Jacobian kernel variants (tiles per warp / blocks per SM): parity subset and bench A/B
mkdir -p gpurun_out
for v in 2 3; do
NBG_JAC_MMA=$v timeout 300 python -m pytest tests -m gpu -x -q -k "trappist8 or full_size or one_shot or transit_parameters" > gpurun_out/pytest_mma$v.log 2>&1; tail -2 gpurun_out/pytest_mma$v.log
done
for v in 0 1 2 3; do
NBG_JAC_MMA=$v timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/bench_mma$v.json 2> gpurun_out/bench_mma$v.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for t in ("mma0", "mma1", "mma2", "mma3"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % t))
        print(t, "value %.4g" % d["value"], {k: round(v) for k, v in d["kernel_ms"].items()})
    except Exception as ex:
        print(t, "failed", ex)
PY
