# Final round script: parity tests, smoke, bench (both arms), serialized bench, ncu launch list + full capture of the dominant kernel
# (-> profiles/r02_jac_rx_profile.json, which bench.py then quotes), SASS listing, the full-length run.  ROUND names the outputs.
mkdir -p gpurun_out
R=${ROUND:-r02z}
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
LIB=nbodygradient.jl_b200/csrc/libnbgrad_b200.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jac_rx_kernel -s 1 -c 1 -f -o /tmp/${R}_jac_rx_kernel python bench.py --steps 1 --warmup 1 --nsys 16384 --window 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_jac_rx_kernel.log 2>&1
python tools/ncu_summary.py /tmp/${R}_jac_rx_kernel.ncu-rep > gpurun_out/${R}_jac_rx_kernel.txt 2>&1
python tools/ncu_hot.py /tmp/${R}_jac_rx_kernel.ncu-rep $LIB jac_rx_kernel 30 2>&1 | cut -c1-220 > gpurun_out/${R}_jac_rx_kernel_hot_lines.txt
python tools/ncu_profile_json.py /tmp/${R}_jac_rx_kernel.ncu-rep 32 gpurun_out/r02_jac_rx_profile.json
cp gpurun_out/r02_jac_rx_profile.json profiles/r02_jac_rx_profile.json
timeout 400 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/bench_ref.err
NBG_OVERLAP=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/${R}_bench_serialized.json 2> gpurun_out/bench_ser.err
timeout 300 python bench.py --no-cpu-baseline --steps 3 --e2e-input elements --e2e-output chi2 > gpurun_out/${R}_bench_elements_chi2.json 2> gpurun_out/bench_chi2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
cuobjdump -sass -fun "$(cuobjdump -elf $LIB 2>/dev/null | grep -o '_ZN[^ ]*jac_rx_kernelILi8ELi4ELb0ELi2ELb0ELb1E[^ ]*' | head -1)" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4,}\*/" | awk '{print $2}' | sed 's/;//' | sort | uniq -c | sort -rn | head -40 > gpurun_out/${R}_jac_rx_sass_mix.txt
timeout 900 python bench.py --full --full-output chi2 > gpurun_out/${R}_bench_full_chi2_65536.json 2> gpurun_out/bench_full.err; echo "full rc=$?"
tail -n 3 gpurun_out/smoke.log; tail -n 3 gpurun_out/bench.err; cat gpurun_out/${R}_bench.json; cat gpurun_out/${R}_bench_reference.json; cat gpurun_out/${R}_bench_full_chi2_65536.json
python - <<'PY'
import json, os
R = os.environ.get("ROUND", "r02z")
d = json.load(open("gpurun_out/%s_bench_serialized.json" % R))
print("serialized", "value %.4g" % d["value"], {k: round(v) for k, v in d["kernel_ms"].items()})
PY
