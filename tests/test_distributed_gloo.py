"""N > 1 host logic on CPU: world_size-2 gloo processes exercise env sharding and the flat gradient-bucket all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from drl_on_robot_arm_b200 import distributed as D
    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                                   # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(9, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    if rank == 1:                                          # desynchronise, then re-sync with one broadcast
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)
    D.broadcast_module(net, src=0)
    bucket = D.GradBucket(net.parameters())
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    torch.manual_seed(100 + rank)                          # different data per rank
    x, y = torch.randn(32, 9), torch.randn(32, 1)
    for _ in range(3):
        bucket.zero()
        loss = torch.nn.functional.mse_loss(net(x), y)
        loss.backward()
        local = bucket.flat.clone()
        bucket.allreduce_mean()
        opt.step()
    lo, hi = D.shard_range(4097, rank, world)
    stats = D.allreduce_scalars({"succ": float(rank + 1), "n": 1.0})
    flat_w = torch.cat([p.data.reshape(-1) for p in net.parameters()])
    class Holder:                                          # stands in for a VectorTrainer holding recorded collectives
        released = 0

        def release_graphs(self):
            Holder.released += 1
    h = Holder()
    D.register_graph_holder(h)
    D.shutdown()                                           # ordered teardown: holders release first, then barrier + destroy
    out.put((rank, flat_w.numpy(), local.numpy(), bucket.flat.numpy().copy(), (lo, hi), stats, bucket.numel,
             Holder.released, dist.is_initialized()))


def test_gloo_world2_bucket_allreduce_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(t[7] == 1 and t[8] is False for t in res)   # distributed.shutdown(): graphs released once, group destroyed
    (r0, w0, l0, g0, s0, st0, n0), (r1, w1, l1, g1, s1, st1, n1) = [t[:7] for t in res]
    assert np.array_equal(w0, w1)                          # replicas stay bit-identical after 3 averaged steps
    assert np.allclose(g0, (l0 + l1) / 2, atol=1e-7) and np.array_equal(g0, g1)   # bucket holds the mean gradient
    assert not np.allclose(l0, l1)
    assert n0 == 9 * 16 + 16 + 16 + 1
    assert s0 == (0, 2049) and s1 == (2049, 4097)
    assert st0 == st1 == {"n": 2.0, "succ": 3.0}


def test_shard_range_partitions():
    from drl_on_robot_arm_b200.distributed import shard_range
    for n in (0, 1, 7, 4096, 32768, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(32768, 3, 8) == (12288, 16384)     # BASELINE config 5: 8 x 4096
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_grad_bucket_single_process_views():
    from drl_on_robot_arm_b200.distributed import GradBucket
    net = torch.nn.Linear(4, 3)
    b = GradBucket(net.parameters())
    assert b.numel == 15 and b.nbytes == 60
    net(torch.ones(2, 4)).sum().backward()
    assert torch.equal(b.flat[:12].view(3, 4), net.weight.grad) and float(b.flat[:12].sum()) == 24.0
    assert b.allreduce_mean() is None
    net.weight.grad = None                                 # e.g. optimizer.zero_grad(set_to_none=True)
    b.zero()
    assert net.weight.grad is not None and net.weight.grad.data_ptr() == b.flat.data_ptr()
    assert float(b.flat.abs().sum()) == 0.0
