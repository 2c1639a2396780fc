"""Entry points and run-level behaviour on the device: the resume artefact continues a run bit-identically (SURVEY.md
section 8 row f4), and code/LA_BCP_train.py delivers the graphed step."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import bcp_oracle as O
from tests.util import record

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fresh(dev, seed=7):
    from bcp_b200.networks.net_factory import net_factory
    from bcp_b200.optim import FusedSGD_EMA
    model, ema = net_factory("VNet", 1, 2, "train"), net_factory("VNet", 1, 2, "train")     # with Dropout3d: RNG matters
    for p in ema.parameters():
        p.detach_()
    O.fill_state_dict_(model, seed)
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    opt = FusedSGD_EMA(model, ema, lr=0.01, momentum=0.9, weight_decay=1e-4, ema_alpha=0.99)
    return model, ema, opt


@pytest.mark.parametrize("graphed", [False, True])
def test_resume_is_bit_identical(tmp_path, graphed):
    """4 steps straight == 2 steps, save_resume, NEW objects, load_resume, 2 steps: weights, EMA teacher, momentum, BN
    buffers and the loss of every step agree bit for bit (np.random box draws, torch CUDA dropout masks and the
    learning-rate change at step 3 included)."""
    from bcp_b200.graph import GraphedStep
    from bcp_b200.step import la_self_train_step
    from bcp_b200.utils.checkpoint import load_resume, save_resume
    dev = torch.device("cuda:0")
    shape = (32, 32, 16)
    batches = [(O.synthetic_volume((8, 1) + shape, 30 + i).to(dev), O.synthetic_labels((8,) + shape, 40 + i).to(torch.uint8).to(dev))
               for i in range(4)]

    def runner(model, ema, opt):
        if not graphed:
            return lambda v, l: la_self_train_step(model, ema, opt, v, l, labeled_bs=4)
        state = (np.random.get_state(), torch.cuda.get_rng_state(dev))
        gs = GraphedStep("la", model, ema, opt, (8, 1) + shape, labeled_bs=4)
        np.random.set_state(state[0])
        torch.cuda.set_rng_state(state[1], dev)
        return lambda v, l: gs(v, l)

    def steps(step, opt, lo, hi, losses):
        for i in range(lo, hi):
            if i == 3:
                opt.param_groups[0]["lr"] = 0.001
            losses.append(float(step(*batches[i])["loss"]))

    np.random.seed(5)
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    model, ema, opt = _fresh(dev)
    ref_losses = []
    steps(runner(model, ema, opt), opt, 0, 4, ref_losses)
    ref = [t.clone() for t in (model.runtime.arena, ema.runtime.arena, opt.buf)] + [b.clone() for b in model.runtime.int_buffers]

    np.random.seed(5)
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    model, ema, opt = _fresh(dev)
    losses = []
    steps(runner(model, ema, opt), opt, 0, 2, losses)
    path = tmp_path / "resume.pth"
    save_resume(path, model, opt, ema, iteration=2, stage="self_train", extra={"best_dice": 0.5})
    np.random.seed(99)                                     # scramble every stream: the artefact must bring them back
    torch.manual_seed(99)
    torch.cuda.manual_seed(99)
    model2, ema2, opt2 = _fresh(dev, seed=123)
    step2 = runner(model2, ema2, opt2)                     # (graph capture happens BEFORE the state is loaded)
    it, stage, extra = load_resume(path, model2, opt2, ema2)
    assert (it, stage, extra["best_dice"]) == (2, "self_train", 0.5)
    assert opt2.step_count == 2
    steps(step2, opt2, 2, 4, losses)
    got = [model2.runtime.arena, ema2.runtime.arena, opt2.buf] + list(model2.runtime.int_buffers)
    assert losses == ref_losses, (losses, ref_losses)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    # the reference's own loader reads the artefact: torch.load(path)['net'] (LA_BCP_train.py:91-93)
    sd = torch.load(str(path), weights_only=False)
    assert set(sd["net"].keys()) == set(model.state_dict().keys()) and "opt" in sd


def test_la_entry_script_runs_graphed(tmp_path):
    """code/LA_BCP_train.py --max_steps 60 (synthetic volumes, both stages): finishes, logs finite losses, writes the
    reference's snapshot files, and its self-training rate is that of the graphed step, not of an eager Python loop
    (eager enqueue alone costs ~10 ms per step; the graphed step with the per-step H2D copy runs > 80 it/s on a B200)."""
    cmd = [sys.executable, os.path.join(ROOT, "code", "LA_BCP_train.py"), "--max_steps", "60", "--log_every", "20", "--ckpt_every", "40"]
    out = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    losses = [float(m.group(1)) for m in re.finditer(r"iteration \d+ : loss: ([0-9.naninf-]+)", out.stdout)]
    assert len(losses) == 6 and all(np.isfinite(losses)), out.stdout[-1500:]
    rates = {m.group(1): float(m.group(2)) for m in re.finditer(r"(pre_train|self_train): 60 iterations, ([0-9.]+) it/s", out.stdout)}
    record("entry_script_self_train_it_per_s", rates.get("self_train", 0.0))
    record("entry_script_pre_train_it_per_s", rates.get("pre_train", 0.0))
    base = tmp_path / "model" / "BCP" / "LA_BCP_8_labeled"
    assert (base / "pre_train" / "VNet_best_model.pth").exists() and (base / "self_train" / "VNet_best_model.pth").exists()
    assert (base / "self_train" / "resume.pth").exists()
    assert rates["self_train"] >= 60.0, rates
