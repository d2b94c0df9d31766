# A/B of two prebuilt libraries (libA.so = HEAD, libB.so = working tree), operator kernels serialized for clean per-kernel times
for L in /root/repo/libA.so /root/repo/libB.so; do
NBG_OVERLAP=${OVL:-1} NBGRAD_B200_LIB=$L timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print('lib=[$L] value %.4g  kernel_ms %s' % (d['value'], {k: round(v,1) for k,v in d['kernel_ms'].items()}))
PY
done
