# usage: VAR=NAME VALS="a b c" bash tools/gpu_env_ab.sh  -- parity tests once, then a short bench per value of the env variable
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for v in $VALS; do
  env $VAR=$v timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_ab_$v.json'))
    print('$VAR=$v value %.4g  kernel_ms %s  newton %.2f itmax %d' % (d['value'], {k: round(x,1) for k,x in d['kernel_ms'].items()}, d['rates']['newton_iters_per_transit'], d['status_bits']['transit_itmax']))
except Exception as e:
    print('$VAR=$v failed', e); print(open('gpurun_out/bench_ab_$v.err').read()[-2000:])
PY
done
