"""N>1 host-side logic on CPU: two gloo ranks shard a batch, reduce the timing (max over ranks), sum work counters and
gather per-rank summaries.  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest


def test_shard_ranges_cover_and_are_disjoint():
    import nbgrad as nb
    for nsys in (1, 2, 7, 64, 65536, 1048576, 1048577):
        for world in (1, 2, 3, 4, 8):
            if nsys < world:
                continue
            rs = [nb.shard_range(nsys, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == nsys
            assert all(rs[r][1] == rs[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == nb.shard_counts(nsys, world)
    with pytest.raises(ValueError):
        nb.shard_range(8, 2, 2)


def _worker(rank, world, port, nsys, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "nbodygradient.jl_b200")]
    import torch.distributed as dist
    import nbgrad as nb
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = nb.shard_range(nsys, rank, world)
    # stand-in for the per-rank device results: a function of the GLOBAL system index only
    local_counts = np.stack([np.arange(lo, hi) * 3 + 1, np.arange(lo, hi) % 5], axis=1).astype(np.int64)
    ms = nb.max_over_ranks(dist, 10.0 + 5.0 * rank)
    tot = nb.sum_over_ranks(dist, [hi - lo, float(local_counts.sum())])
    full = nb.gather_slices(dist, local_counts, nsys)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ms, tot.tolist(), full.tolist()))


@pytest.mark.parametrize("nsys", [8, 11])
def test_two_rank_gloo_shard_reduce_gather(nsys):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, nsys, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    expect = np.stack([np.arange(nsys) * 3 + 1, np.arange(nsys) % 5], axis=1)
    for rank, ms, tot, full in res:
        assert ms == 15.0                                  # max over ranks, identical on every rank
        assert tot == [float(nsys), float(expect.sum())]   # every system counted exactly once
        assert np.array_equal(np.array(full), expect)      # gather restores the global order


def test_reference_arm_under_torchrun_prints_one_line():
    # the driver launches `bench.py --impl reference` like the GPU arm (torchrun for N > 1): rank 0 alone times the CPU oracle and
    # prints ONE JSON line with the contract's keys, the other ranks exit 0 without work.  No GPU involved.
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--ref-window", "16",
           "--ref-systems-per-core", "1"]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "system-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "cfg 5" in d["config"]["workload"] and d["config"]["batch_per_gpu"] == 131072
