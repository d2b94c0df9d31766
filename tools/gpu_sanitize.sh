# compute-sanitizer on the TransitParameters / ti = 1 test (flaky across builds: passes or gives a fixed wrong v_sky gradient)
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
for i in 1 2 3; do timeout 300 python -m pytest tests -m gpu -q -k "transit_parameters_and_ti" 2>&1 | tail -n 2; done
for tool in memcheck racecheck initcheck synccheck; do
  echo "=== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests -m gpu -q -x -k "transit_parameters_and_ti" > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|Uninitialized|hazard|Barrier" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
  grep -B2 -A12 "=========     at " gpurun_out/sanitize_$tool.log | head -60
done
