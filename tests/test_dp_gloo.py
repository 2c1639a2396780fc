"""World-size-2 data-parallel host logic on CPU (gloo): the flat-gradient all-reduce used by
bcp_b200.optim (one collective per step, 1/world folded into the update) reproduces the single-process step on the
union of the per-rank batches with per-rank BatchNorm statistics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bcp_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tiny_net():
    net = O.OracleVNet(1, 2, 2, "batchnorm", False)
    O.fill_state_dict_(net, 5)
    return net


def _rank_batch(rank):
    x = O.synthetic_volume((2, 1, 48, 48, 48), 100 + rank)
    y = O.synthetic_labels((2, 48, 48, 48), 200 + rank)
    return x, y


def _local_grads(net, x, y):
    net.zero_grad()
    out, _ = net(x)
    loss = O.mix_loss_la(out, y, y, torch.ones(2, 48, 48, 48, dtype=torch.int64))
    loss.backward()
    train = [p for n, p in net.named_parameters() if n.startswith(("encoder.", "decoder."))]
    return torch.cat([p.grad.reshape(-1) for p in train]), float(loss)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bcp_b200.optim import _FusedBase

        class _RT:            # the part of NetRuntime the all-reduce touches
            pass
        net = _tiny_net()
        net.train()
        flat, loss = _local_grads(net, *_rank_batch(rank))
        fb = _FusedBase.__new__(_FusedBase)
        fb.rt = _RT()
        fb.rt.grad_arena = flat.clone()
        fb.dev = flat.device             # no weight-gradient side stream pending on the CPU
        fb._reduced_upto = None          # no bucket in flight: the whole arena goes in one collective
        w = fb._allreduce()
        q.put((rank, w, fb.rt.grad_arena.numpy().copy(), loss))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_grad_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == 2 and res[1][1] == 2
    assert np.array_equal(res[0][2], res[1][2])                     # replicas see identical reduced gradients
    # single-process reference: sum of the two per-rank gradients (each rank = its own BN batch)
    net = _tiny_net()
    net.train()
    g0, _ = _local_grads(net, *_rank_batch(0))
    net = _tiny_net()
    net.train()
    g1, _ = _local_grads(net, *_rank_batch(1))
    np.testing.assert_allclose(res[0][2], (g0 + g1).numpy(), rtol=1e-5, atol=1e-7)
    # folding 1/world into the update == SGD on the mean gradient
    p = torch.randn(g0.numel())
    lr, world_scale = 0.01, 1.0 / world
    np.testing.assert_allclose((p - lr * (torch.from_numpy(res[0][2]) * world_scale)).numpy(),
                               (p - lr * (g0 + g1) / 2).numpy(), rtol=1e-6, atol=1e-7)


def test_rank_sharding_is_disjoint():
    import bench
    a, la = bench.synthetic_batch("la", 0)
    b, lb = bench.synthetic_batch("la", 1)
    assert a.shape == (8, 1, 112, 112, 80) and la.dtype == torch.uint8
    assert not torch.equal(a, b)
    assert 0.01 < float(la.float().mean()) < 0.5
