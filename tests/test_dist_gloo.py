"""CPU: the N>1 plumbing of bench.py with world_size=2 over gloo (127.0.0.1): barrier, max-over-ranks of
the per-rank device time, whole-job throughput arithmetic, and the reference arm's rank gating."""
import os
import socket
import subprocess
import sys

import torch
import torch.multiprocessing as mp

from tests.util import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import bench
    r, w, l = bench.dist_setup(world)
    assert (r, w, l) == (rank, world, rank)
    bench.barrier(w)
    ms = bench.max_over_ranks(10.0 + 5.0 * rank, w, torch.device("cpu"))  # rank 1 is the slow one
    value = w * bench.PER_GPU_BATCH * 4 / (ms * 1e-3)
    q.put((rank, ms, value))
    bench.barrier(w)
    import torch.distributed as dist
    dist.destroy_process_group()


def test_max_over_ranks_and_weak_scaling_value_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [15.0, 15.0]  # both ranks agree on the max
    assert abs(res[0][2] - 2 * 2 * 4 / 0.015) < 1e-6  # whole-job images / slowest rank's time


def test_reference_arm_nonzero_rank_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
