"""Layer-by-layer comparison of the sm_100a V-Net against the fp32 oracle on the same GPU (train mode)."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import bcp_oracle as O
from tests.golden.golden_common import inject_dropout
from tests.util import planar_from_cb8, rel_rms
from bcp_b200.networks.VNet import VNet, _Stage3d

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def run(shape, has_dropout, train=True, seed=23):
    x = O.synthetic_volume(shape, 77).to(dev)
    net = VNet(1, 2, 16, "batchnorm", has_dropout)
    O.fill_state_dict_(net, seed)
    net = net.to(dev).train(train)
    ref = O.OracleVNet(1, 2, 16, "batchnorm", has_dropout)
    O.fill_state_dict_(ref, seed)
    ref = ref.to(dev).train(train)
    if has_dropout:
        inject_dropout(net, seed=24)
        inject_dropout(ref, seed=24)
    mine, theirs = {}, {}
    for name, m in net.named_modules():
        if isinstance(m, _Stage3d):
            m.register_forward_hook(lambda mod, inp, out, name=name: mine.__setitem__(name, out.detach()))
    for name, m in ref.named_modules():
        if isinstance(m, O._Wrap):
            m.register_forward_hook(lambda mod, inp, out, name=name: theirs.__setitem__(name, out.detach().clone()))
    with torch.no_grad():
        lo, _ = net(x, with_features=False)
        lr, _ = ref(x)
    print(f"shape={shape} dropout={has_dropout} train={train}: logits rel_rms {rel_rms(lo, lr):.4f}")
    skips = {"decoder.block_five_up": "encoder.block_four", "decoder.block_six_up": "encoder.block_three",
             "decoder.block_seven_up": "encoder.block_two", "decoder.block_eight_up": "encoder.block_one"}
    for name in mine:
        c = mine[name].shape[1] * 8
        a = planar_from_cb8(mine[name], c)
        b = theirs[name]
        if name in skips:
            b = b + theirs[skips[name]]
        print(f"   {name:28s} rel_rms {rel_rms(a, b):.4f}   ref rms {float(b.pow(2).mean().sqrt()):.3f}  zero-frac mine {float((a == 0).float().mean()):.3f} ref {float((b == 0).float().mean()):.3f}")


def autocast_baseline(shape, seed=23):
    """How far is stock PyTorch bf16 autocast (cuDNN) from fp32 on the same fixture?  Calibrates the bf16 budget."""
    x = O.synthetic_volume(shape, 77).to(dev)
    w = O.synthetic_volume((shape[0], 2) + shape[2:], 78).to(dev)
    outs = {}
    for mode in ("fp32", "bf16"):
        ref = O.OracleVNet(1, 2, 16, "batchnorm", False)
        O.fill_state_dict_(ref, seed)
        ref = ref.to(dev).train()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            lo, _ = ref(x)
        (lo.float() * w).sum().backward()
        outs[mode] = (lo.detach().float(), {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None})
    print(f"[autocast bf16 vs fp32] shape={shape} logits rel_rms {rel_rms(outs['bf16'][0], outs['fp32'][0]):.4f}")
    for n in ("decoder.out_conv.weight", "decoder.block_nine.conv.0.weight", "decoder.block_seven.conv.0.weight",
              "encoder.block_five.conv.0.weight", "encoder.block_three.conv.0.weight", "encoder.block_one.conv.0.weight"):
        print(f"      grad {n:40s} rel_rms {rel_rms(outs['bf16'][1][n], outs['fp32'][1][n]):.4f}")


if __name__ == "__main__":
    autocast_baseline((2, 1, 112, 112, 80))
    autocast_baseline((2, 1, 48, 48, 48))
    sys.exit(0)
    run((2, 1, 112, 112, 80), False, True)
    run((2, 1, 112, 112, 80), False, False)
    run((2, 1, 112, 112, 80), True, True)
    run((2, 1, 48, 48, 48), False, True)
