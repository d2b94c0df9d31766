# r01l: DMMA Jacobian kernel: parity (N = 8 tests) and bench A/B; e2e with balanced chunks
mkdir -p gpurun_out
NBG_JAC_MMA=1 timeout 600 python -m pytest tests -m gpu -x -q -k "trappist8 or step_parity or nbody_sweep or full_size or one_shot or ntt_overflow or small_event or resume or transit_parameters or cfg2_full" > gpurun_out/pytest_mma.log 2>&1; tail -8 gpurun_out/pytest_mma.log
NBG_JAC_MMA=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_mma.json 2> gpurun_out/bench_mma.err; echo "bench rc=$?"
NBG_OUT_SLICES=8 timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_rx.json 2> gpurun_out/bench_rx.err; echo "bench rc=$?"
python - <<'PY'
import json
for t in ("mma", "rx"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % t))
        print(t, "value %.4g e2e %.4g chunk %d" % (d["value"], d["e2e"]["value"], d["config"]["chunk_steps"]), {k: round(v) for k, v in d["kernel_ms"].items()})
    except Exception as ex:
        print(t, "failed", ex)
PY
tail -3 gpurun_out/bench_mma.err
