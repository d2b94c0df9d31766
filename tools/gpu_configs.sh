# the other BASELINE configurations on one GPU: cfg 3 (outer solar system, grad = false, 10^6 steps), cfg 4 (N = 2..16, grad = true, exactly 10^4 steps)
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
timeout 3000 python tools/bench_configs.py --cfg3-steps ${CFG3_STEPS:-1000000} --nmin ${NMIN:-2} --nmax ${NMAX:-16} > gpurun_out/r02_configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
tail -n 3 gpurun_out/configs.err; cut -c1-400 gpurun_out/r02_configs.jsonl
