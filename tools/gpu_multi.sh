# multi-GPU pass (gpurun --gpus NG): the multi-device plan inside the library (tests + single-process bench) and the torchrun arm
NG=${NG:-2}
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests -m gpu -q -k "multi_device" -s > gpurun_out/pytest_multi_${NG}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_${NG}gpu.log
tail -n 4 gpurun_out/pytest_multi_${NG}gpu.log
timeout 600 python bench.py --gpus $NG --single-process --steps 3 --warmup 3 > gpurun_out/r02_bench_${NG}gpu_single_process.json 2> gpurun_out/bench_sp.err; echo "single-process rc=$?"
tail -n 2 gpurun_out/bench_sp.err; cat gpurun_out/r02_bench_${NG}gpu_single_process.json
timeout 600 python bench.py --gpus $NG --single-process --steps 3 --warmup 3 --e2e-output chi2 --no-cpu-baseline > gpurun_out/r02_bench_${NG}gpu_single_process_chi2.json 2> gpurun_out/bench_sp2.err; echo "single-process chi2 rc=$?"
tail -n 2 gpurun_out/bench_sp2.err; cat gpurun_out/r02_bench_${NG}gpu_single_process_chi2.json
if [ "${TORCHRUN:-1}" = "1" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${NG}gpu_torchrun.json 2> gpurun_out/bench_tr.err; echo "torchrun rc=$?"
tail -n 2 gpurun_out/bench_tr.err; cat gpurun_out/r02_bench_${NG}gpu_torchrun.json
fi
