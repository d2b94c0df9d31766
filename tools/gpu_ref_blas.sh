#!/bin/bash
# the CPU reference arm on the GPU box's host cores: dense products through OpenBLAS (as Julia's mul!) against the built-in loop nests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
make -C oracle -s
nproc; lscpu | grep "Model name"
for i in 1 2; do
  timeout 100 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02z_ref_blas_$i.json
  NBGRAD_REF_NO_BLAS=1 timeout 100 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02z_ref_loops_$i.json
done
python - <<'PY'
import json
for k in ("blas_1", "loops_1", "blas_2", "loops_2"):
    d = json.load(open("gpurun_out/r02z_ref_%s.json" % k))
    print(k, round(d["value"]), d["cpu_baseline"]["cores"], d["cpu_baseline"]["sample"][:120])
PY
