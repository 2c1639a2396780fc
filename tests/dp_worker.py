"""torchrun worker of tests/test_multigpu_dp.py: two data-parallel LA self-training steps per rank (NCCL), results to <out>/rank<r>.pt.

usage: python -m torch.distributed.run --nproc-per-node N ... tests/dp_worker.py <out_dir> <graphed 0|1> <overlap 0|1>"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
SHAPE = (32, 32, 16)
STEPS = 2


def build(dev, seed=7):
    from oracle import bcp_oracle as O
    from bcp_b200.networks.net_factory import net_factory
    from bcp_b200.optim import FusedSGD_EMA
    model, ema = net_factory("VNet", 1, 2, "test"), net_factory("VNet", 1, 2, "test")     # no Dropout3d: one RNG stream less
    for p in ema.parameters():
        p.detach_()
    O.fill_state_dict_(model, seed)
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    return model, ema, FusedSGD_EMA(model, ema, lr=0.01, momentum=0.9, weight_decay=1e-4, ema_alpha=0.99)


def rank_batch(rank, step, dev):
    from oracle import bcp_oracle as O
    vol = O.synthetic_volume((8, 1) + SHAPE, 1000 + 10 * rank + step).to(dev)
    lab = O.synthetic_labels((8,) + SHAPE, 2000 + 10 * rank + step).to(torch.uint8).to(dev)
    return vol, lab


def main():
    out_dir, graphed, overlap = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    os.environ["BCP_DP_OVERLAP"] = str(overlap)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from bcp_b200.step import la_self_train_step
    from bcp_b200.graph import GraphedStep
    # rank-dependent construction seed on purpose: the optimiser's start-up broadcast must make the replicas identical
    model, ema, opt = build(dev, seed=7 + rank)
    np.random.seed(100 + rank)
    gs = GraphedStep("la", model, ema, opt, (8, 1) + SHAPE, labeled_bs=4) if graphed else None
    np.random.seed(100 + rank)
    losses = []
    for s in range(STEPS):
        vol, lab = rank_batch(rank, s, dev)
        r = gs(vol, lab) if gs is not None else la_self_train_step(model, ema, opt, vol, lab, labeled_bs=4)
        losses.append(float(r["loss"]))
    torch.cuda.synchronize()
    rt, ert = model.runtime, ema.runtime
    torch.save({"params": rt.arena[:rt.n_param].cpu(), "ema_params": ert.arena[:ert.n_param].cpu(), "momentum": opt.buf.cpu(),
                "losses": losses, "world": world}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    if gs is not None:                 # a captured collective must be released before its communicator is torn down
        gs.graph.reset()
        del gs.graph
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
