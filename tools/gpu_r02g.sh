# round 2, GPU pass g: pivot-block variants of the N = 8 Jacobian kernel (U = 8 production, 4, 2), block test
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -q -k "block_scaled or transit_parameters or fused_chi2 or step_parity" -s > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g.log
grep -E "passed|failed|FAILED|blocks|rc=|trial" gpurun_out/pytest_gpu_g.log | cut -c1-300 | tail -14
for lib in nbodygradient.jl_b200/csrc/libnbgrad_b200.so ab/libnbg_u4.so ab/libnbg_u2.so nbodygradient.jl_b200/csrc/libnbgrad_b200.so ab/libnbg_u4.so ab/libnbg_u2.so; do
  timeout 200 python tools/ab_time.py $lib 4 2>&1 | tail -n 1 | tee -a gpurun_out/r02g_ab.jsonl
done
