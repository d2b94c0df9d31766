# A/B of run-time variants (environment knobs) on one box: short device-resident bench runs, then a parity subset
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; }
run base NBG_PHI_CACHED=1
run cached NBG_PHI_CACHED=2
run bases NBG_PHI_CACHED=1 NBG_OVERLAP=0
run cacheds NBG_PHI_CACHED=2 NBG_OVERLAP=0
NBG_PHI_CACHED=2 timeout 400 python -m pytest tests -m gpu -x -q -k "not cfg2_full and not full_size" 2>&1 | tail -3
python - <<'PY'
import json
for t in ("base", "cached", "bases", "cacheds"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % t))
        print(t, "value %.4g" % d["value"], {k: round(v) for k, v in d["kernel_ms"].items()})
    except Exception as ex:
        print(t, "failed", ex)
PY
