# A/B of run-time variants (environment knobs) on one box: short device-resident bench runs.  usage: edit the `run` lines.
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; }
run pre2 NBG_NEWTON_PRE=2
run pre3 NBG_NEWTON_PRE=3
run pre4 NBG_NEWTON_PRE=4
python - <<'PY'
import json
for t in ("pre2", "pre3", "pre4"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % t))
        print(t, "value %.4g" % d["value"], {k: round(v) for k, v in d["kernel_ms"].items()}, d["rates"]["newton_iters_per_transit"], d["status_bits"])
    except Exception as ex:
        print(t, "failed", ex)
PY
