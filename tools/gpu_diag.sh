mkdir -p gpurun_out
PYTHONPATH=nbodygradient.jl_b200 python -m nbgrad.build --if-stale 2>&1 | tail -1
python tools/diag_tp.py 1 2>&1 | tail -n 30
NBG_FORCE_GENERIC_JAC=1 python tools/diag_tp.py 1 2>&1 | tail -n 12
NBG_OVERLAP=0 python tools/diag_tp.py 1 2>&1 | tail -n 8
NBG_NEWTON_PRE=0 python tools/diag_tp.py 1 2>&1 | tail -n 8
python tools/diag_tp.py 0 2>&1 | tail -n 6
