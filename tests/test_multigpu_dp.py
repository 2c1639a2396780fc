"""N-rank data parallelism on real GPUs over NCCL: after two steps on different per-rank
batches the replicas' parameters, EMA-teacher parameters and momentum are bit-identical, and equal to a single-process
emulation that sums the per-rank gradients by hand.  Marker ``multigpu``: run with
``gpurun --gpus 2 -- python -m pytest tests -m multigpu`` (log of the last run: profiles/multigpu_r02.log)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.multigpu       # NOT part of -m gpu: a single-GPU box cannot run it (NCCL refuses two ranks on one device)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _emulate(world, dev):
    """The same two steps in ONE process: replica r runs forward/backward on its own batch and box stream, the gradient arenas
    are summed in rank order and every replica applies the sum with grad_scale 1/world."""
    from bcp_b200 import optim as OPT
    from bcp_b200.step import la_self_train_step
    from tests import dp_worker as W
    reps = [W.build(dev, seed=7) for _ in range(world)]              # the broadcast makes every replica rank 0's
    streams = []
    for r in range(world):
        np.random.seed(100 + r)
        streams.append(np.random.get_state())

    class _FakeDist:
        @staticmethod
        def all_reduce(t, *a, **k):
            return None

        @staticmethod
        def broadcast(t, *a, **k):
            return None
    real_world = OPT._world
    OPT._world = lambda: (_FakeDist, world)
    try:
        for _, _, opt in reps:
            opt._hyper_host = None
            opt.refresh_hyper()                                           # grad_scale = 1 / world
        for s in range(W.STEPS):
            grads = []
            for r, (model, ema, opt) in enumerate(reps):
                np.random.set_state(streams[r])
                real_step = opt.step
                opt.step = lambda: None                                   # forward/backward only
                la_self_train_step(model, ema, opt, *W.rank_batch(r, s, dev), labeled_bs=4)
                opt.step = real_step
                streams[r] = np.random.get_state()
                grads.append(model.runtime.grad_arena.clone())
            total = grads[0].clone()
            for g in grads[1:]:
                total += g                                                # rank order, like a ring of two
            for model, ema, opt in reps:
                model.runtime.grad_arena.copy_(total)
                opt.step()
    finally:
        OPT._world = real_world
    model, ema, opt = reps[0]
    rt, ert = model.runtime, ema.runtime
    return rt.arena[:rt.n_param].cpu(), ert.arena[:ert.n_param].cpu(), opt.buf.cpu()


@pytest.mark.parametrize("graphed,overlap", [(0, 0), (1, 0), (1, 1)])
def test_two_rank_step_matches_single_process(tmp_path, graphed, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py"), str(tmp_path), str(graphed), str(overlap)]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]          # includes a clean destroy_process_group()
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    for k in ("params", "ema_params", "momentum"):
        assert torch.equal(r0[k], r1[k]), k                                       # replicas stay bit-identical
    assert r0["losses"] != r1["losses"]                                           # ... on different data
    p, e, m = _emulate(2, torch.device("cuda:0"))
    assert torch.equal(r0["params"], p) and torch.equal(r0["ema_params"], e) and torch.equal(r0["momentum"], m)
