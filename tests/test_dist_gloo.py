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


def _reducer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import torch.nn as nn
    from pairnet_b200.trainer import GradReducer
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                     # identical replicas
    net = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 4))
    params = list(net.parameters())
    red = GradReducer(params, bucket_bytes=300)              # three small buckets
    opt = torch.optim.SGD(params, lr=0.1)
    x = torch.full((3, 8), float(rank + 1))                  # per-rank data
    for step in range(2):
        net(x).square().sum().backward()
        n = red.finish()
        assert n == len(red.buckets) and n >= 3              # every bucket was reduced from its hook
        opt.step()
        red.zero_grad()
    # the same two steps on the concatenated batch, single process, mean of the per-rank gradients
    torch.manual_seed(0)
    ref = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 4))
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for step in range(2):
        ropt.zero_grad()
        sum(ref(torch.full((3, 8), float(r + 1))).square().sum() for r in range(world)).div(world).backward()
        ropt.step()
    err = max(float((a - b).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
    q.put((rank, err, [float(p.sum()) for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_gloo():
    """GradReducer (trainer.py): gradients accumulate into flat buckets, each bucket is all-reduced from the autograd
    hook of its last parameter, replicas stay identical and equal the single-process large-batch step."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] < 1e-6 and res[1][1] < 1e-6
    assert res[0][2] == res[1][2]                             # replicas bit-identical after the reduced steps
