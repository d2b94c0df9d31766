#!/bin/bash
# cfg 4 at small N (no transits): dense phisalpha operators on a third stream (NBG_OVERLAP=2) against the default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
for ov in 1 2 1 2; do
  NBG_OVERLAP=$ov timeout 600 python tools/bench_configs.py --skip-cfg3 --nmin 2 --nmax 8 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d.get('config') == 'cfg4':
        print(json.dumps({'overlap': $ov, 'nbody': d['nbody'], 'device_ms': d['device_ms'], 'system_steps_per_s': d['system_steps_per_s'], 'frac_fp64_peak': d['frac_fp64_peak'], 'kernel_ms': d['kernel_ms']}))
" | tee -a gpurun_out/r02z_overlap3_cfg4.jsonl
done
