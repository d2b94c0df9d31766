# round 2, GPU pass j: cfg 3 long-run scaling (A/B against the r01 build); N = 15, 16 register-resident single-buffer kernel (tests + timing)
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
for lib in nbodygradient.jl_b200/csrc/libnbgrad_b200.so ab/libnbg_r01.so; do
  timeout 300 python tools/ab_cfg3.py $lib 300000 2>&1 | tail -n 1 | tee -a gpurun_out/r02j_ab_cfg3.jsonl
done
timeout 900 python -m pytest tests -m gpu -q -k "nbody_sweep or above_8_bodies" -s > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_j.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu_j.log | cut -c1-300 | tail -8
timeout 900 python tools/bench_configs.py --skip-cfg3 --nmin 13 --nmax 16 --cfg4-steps 512 > gpurun_out/r02j_configs_n13_16.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
tail -n 3 gpurun_out/configs.err; cut -c1-330 gpurun_out/r02j_configs_n13_16.jsonl
NBG_FORCE_GENERIC_JAC=1 timeout 900 python tools/bench_configs.py --skip-cfg3 --nmin 15 --nmax 16 --cfg4-steps 256 2>> gpurun_out/configs.err | cut -c1-330 | tee -a gpurun_out/r02j_configs_n13_16.jsonl
