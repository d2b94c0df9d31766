mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "full_size" > gpurun_out/pytest_full.log 2>&1; tail -5 gpurun_out/pytest_full.log
for K in 1 2 4; do
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-slices $K > gpurun_out/bench_e2e$K.json 2> gpurun_out/bench_e2e$K.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_e2e$K.json')); print('K=$K value %.4g e2e %s' % (d['value'], d['e2e']))
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_e2e$K.err').read()[-1500:])
PY
done
