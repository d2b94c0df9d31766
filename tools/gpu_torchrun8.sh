#!/bin/bash
# the driver's scaling launch at N GPUs: one rank per GPU under torchrun, exactly as the contract states it (cpu baseline only on rank 0)
NG=${NG:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
nproc > gpurun_out/nproc_${NG}gpu.txt; free -g | head -2 >> gpurun_out/nproc_${NG}gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/r02z_bench_${NG}gpu_torchrun.json 2> gpurun_out/bench_tr.err; echo "torchrun rc=$?"
tail -n 3 gpurun_out/bench_tr.err; cat gpurun_out/r02z_bench_${NG}gpu_torchrun.json | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $NG --steps 1 --warmup 1 > gpurun_out/r02z_bench_${NG}gpu_reference.json 2> gpurun_out/bench_tr_ref.err; echo "reference rc=$?"
cat gpurun_out/r02z_bench_${NG}gpu_reference.json | cut -c1-600
