# quick GPU check: parity tests, then short bench runs (device-resident arm only) for each NBG_RX_UNROLL
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for u in ${UNROLLS:-2 1 4}; do
  NBG_RX_UNROLL=$u timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_u$u.json 2> gpurun_out/bench_u$u.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_u$u.json'))
    print('U=$u value %.4g  kernel_ms %s  frac %.3f' % (d['value'], {k: round(v,1) for k,v in d['kernel_ms'].items()}, d['roofline']['frac']))
except Exception as e:
    print('U=$u failed', e); print(open('gpurun_out/bench_u$u.err').read()[-2000:])
PY
done
