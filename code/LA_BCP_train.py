#!/usr/bin/env python
"""LA (left atrium) BCP training entry point on the B200-native kernels.

Keeps the CLI flags, defaults, snapshot directory layout and the two-stage schedule of the reference's
``code/LA_BCP_train.py`` (/root/reference/code/LA_BCP_train.py:32-55,351-371): stage 1 ``pre_train`` (copy-paste
between two labeled volumes), stage 2 ``self_train`` (EMA teacher pseudo-labels, bidirectional copy-paste, SGD, EMA).
The step bodies live in ``bcp_b200.step``.  Extra flags: ``--synthetic`` (seeded synthetic volumes of the LA shape;
the real LA h5 data and h5py are not available in this environment), ``--max_steps`` (bound both stages),
data-parallel launch via ``torchrun`` (one process per GPU, one NCCL all-reduce of the flat gradient per step).

Every step is one CUDA-graph replay (``--graph 1``, bcp_b200/graph.py): the learning-rate decay reaches the device through
the optimiser's hyper vector.  ``--synthetic 0`` reads the real LA set into HBM once (bcp_b200/dataloaders/dataset.py:
h5 via h5py where available, else <name>.npz) and validates every 200 iterations with the sliding-window kernels;
``--ckpt_every N`` / ``--resume 1`` write / continue from one resume artefact per stage (bcp_b200/utils/checkpoint.py).
Out of scope: tensorboard scalars and snapshot images.
"""
import argparse
import logging
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

parser = argparse.ArgumentParser()
parser.add_argument('--root_path', type=str, default='/data/byh_data/SSNet_data/LA', help='Name of Dataset')
parser.add_argument('--exp', type=str, default='BCP', help='exp_name')
parser.add_argument('--model', type=str, default='VNet', help='model_name')
parser.add_argument('--pre_max_iteration', type=int, default=2000, help='maximum pre-train iteration to train')
parser.add_argument('--self_max_iteration', type=int, default=15000, help='maximum self-train iteration to train')
parser.add_argument('--max_samples', type=int, default=80, help='maximum samples to train')
parser.add_argument('--labeled_bs', type=int, default=4, help='batch_size of labeled data per gpu')
parser.add_argument('--batch_size', type=int, default=8, help='batch_size per gpu')
parser.add_argument('--base_lr', type=float, default=0.01, help='maximum epoch number to train')
parser.add_argument('--deterministic', type=int, default=1, help='whether use deterministic training')
parser.add_argument('--labelnum', type=int, default=8, help='trained samples')
parser.add_argument('--gpu', type=str, default='1', help='GPU to use')
parser.add_argument('--seed', type=int, default=1337, help='random seed')
parser.add_argument('--consistency', type=float, default=1.0, help='consistency')
parser.add_argument('--consistency_rampup', type=float, default=40.0, help='consistency_rampup')
parser.add_argument('--magnitude', type=float, default=10.0, help='magnitude')
parser.add_argument('--u_weight', type=float, default=0.5, help='weight of unlabeled pixels')
parser.add_argument('--mask_ratio', type=float, default=2 / 3, help='ratio of mask/image')
parser.add_argument('--u_alpha', type=float, default=2.0, help='unlabeled image ratio of mixuped image')
parser.add_argument('--loss_weight', type=float, default=0.5, help='loss weight of unimage term')
# -- additions
parser.add_argument('--synthetic', type=int, default=1, help='use seeded synthetic LA-shaped volumes')
parser.add_argument('--max_steps', type=int, default=0, help='if > 0, bound the iterations of each stage')
parser.add_argument('--log_every', type=int, default=10)
parser.add_argument('--graph', type=int, default=1, help='replay each step as one CUDA graph (0: eager step functions)')
parser.add_argument('--resume', type=int, default=0, help='continue a stage from <snapshot>/resume.pth when it exists')
parser.add_argument('--ckpt_every', type=int, default=0, help='write the resume artefact every N iterations (0: never)')

patch_size = (112, 112, 80)
num_classes = 2


class SyntheticLA:
    """Seeded stand-in for LAHeart + RandomRotFlip/RandomCrop/ToTensor + TwoStreamBatchSampler
    (dataloaders/dataset.py:91-126,280-307): yields {'image': [B,1,112,112,80] fp32, 'label': [B,112,112,80] uint8},
    labeled samples first."""

    def __init__(self, batch_size, seed, device, pool=4):
        self.bs, self.dev = batch_size, device
        gen = torch.Generator(device="cpu").manual_seed(seed)
        self.pool = []                      # a few pinned host batches, cycled: the per-step H2D copy stays, the host RNG does not
        for _ in range(pool):
            img = torch.randn((self.bs, 1) + patch_size, generator=gen)
            noise = torch.randn((self.bs, 1) + patch_size, generator=gen)
            sm = torch.nn.functional.avg_pool3d(noise, 5, stride=1, padding=2)[:, 0]
            lab = (sm > sm.std()).to(torch.uint8)
            self.pool.append((img.pin_memory(), lab.pin_memory()))

    def __iter__(self):
        k = 0
        while True:
            img, lab = self.pool[k % len(self.pool)]
            k += 1
            yield {"image": img, "label": lab}


class RealLA:
    """LAHeart + Compose([RandomRotFlip, RandomCrop(patch), ToTensor]) + TwoStreamBatchSampler of the reference
    (LA_BCP_train.py:119-133), device-resident (bcp_b200/dataloaders/dataset.py): epochs of the sampler, for ever."""

    def __init__(self, args, device):
        from bcp_b200.dataloaders.dataset import LAHeart, TwoStreamBatchSampler, TwoStreamLoader
        db = LAHeart(base_dir=args.root_path, split='train', num=args.max_samples, device=device)
        labeled_idxs = list(range(args.labelnum))
        unlabeled_idxs = list(range(args.labelnum, min(args.max_samples, len(db))))
        sampler = TwoStreamBatchSampler(labeled_idxs, unlabeled_idxs, args.batch_size, args.batch_size - args.labeled_bs)
        self.loader = TwoStreamLoader(db, sampler, patch_size)
        logging.info("{} iterations per epoch".format(len(self.loader)))

    def __iter__(self):
        while True:
            for batch in self.loader:
                yield batch


def make_loader(args, device, rank):
    if args.synthetic:
        return SyntheticLA(args.batch_size, args.seed + rank, device)
    return RealLA(args, device)


def save_net_opt(net, optimizer, path):
    torch.save({'net': net.state_dict(), 'opt': optimizer.state_dict()}, str(path))      # LA_BCP_train.py:79-84


def load_net(net, path):
    net.load_state_dict(torch.load(str(path), weights_only=False)['net'])                 # LA_BCP_train.py:91-93


def validate(args, model):
    """var_all_case_LA every 200 iterations (LA_BCP_train.py:172-186); needs the real test volumes."""
    if args.synthetic:
        return None
    from bcp_b200.utils import test_3d_patch
    return test_3d_patch.var_all_case_LA(model, num_classes=num_classes, patch_size=patch_size, stride_xy=18, stride_z=4,
                                         root_path=args.root_path)


def run_stage(args, stage, model, ema_model, optimizer, snapshot_path, device, rank, max_iterations):
    """One stage's loop: the whole step is ONE CUDA-graph replay (bcp_b200/graph.py); the host only stages the next batch,
    draws the box, adjusts the learning rate (device hyper vector) and logs.  --graph 0 runs the eager step functions."""
    from bcp_b200.graph import GraphedStep
    from bcp_b200.step import la_pre_train_step, la_self_train_step
    from bcp_b200.utils.checkpoint import load_resume, save_resume
    kind = "la_pre" if stage == "pre_train" else "la"
    iters = max_iterations if not args.max_steps else min(args.max_steps, max_iterations)
    resume_path = os.path.join(snapshot_path, "resume.pth")
    it, best_dice = 0, 0.0
    if args.resume and os.path.exists(resume_path):
        it, st, extra = load_resume(resume_path, model, optimizer, ema_model)
        best_dice = float(extra.get("best_dice", 0.0))
        logging.info("resumed %s at iteration %d from %s" % (stage, it, resume_path))
    gs = None
    if args.graph:
        kw = dict(labeled_bs=args.labeled_bs, mask_ratio=args.mask_ratio)
        if kind == "la":
            kw["u_weight"] = args.u_weight
        state = np.random.get_state()            # capture warm-up draws boxes: keep the stream where the reference would be
        gs = GraphedStep(kind, model, ema_model, optimizer, (args.batch_size, 1) + patch_size, device=device, **kw)
        np.random.set_state(state)
    loader = make_loader(args, device, rank)
    torch.cuda.synchronize(device)
    t0, it0, t_io = time.time(), it, 0.0          # t_io: validation + checkpoint time, excluded from the reported step rate
    for batch in loader:
        if it >= iters:
            break
        if gs is not None:
            r = gs(batch['image'], batch['label'])               # pinned host or device tensors: staged on the copy stream
        elif kind == "la_pre":
            batch = {k: v.to(device, non_blocking=True) for k, v in batch.items() if torch.is_tensor(v)}
            r = la_pre_train_step(model, optimizer, batch['image'], batch['label'], args.labeled_bs, args.mask_ratio)
        else:
            batch = {k: v.to(device, non_blocking=True) for k, v in batch.items() if torch.is_tensor(v)}
            r = la_self_train_step(model, ema_model, optimizer, batch['image'], batch['label'], args.labeled_bs, args.mask_ratio, args.u_weight)
        it += 1
        if it % args.log_every == 0 and rank == 0:
            names = ('loss', 'loss_dice', 'loss_ce') if kind == "la_pre" else ('loss', 'loss_l', 'loss_u')
            vals = tuple(float(r[k]) for k in names)                     # the only device sync of the loop
            rate = (it - it0) / max(time.time() - t0 - t_io, 1e-9)
            logging.info('iteration %d : %s: %03f, %s: %03f, %s: %03f  (%.1f it/s, %.0f patches/s)' %
                         (it, names[0], vals[0], names[1], vals[1], names[2], vals[2], rate, rate * args.labeled_bs))
        if kind == "la" and it % 2500 == 0:                                               # LA_BCP_train.py:273-276
            optimizer.param_groups[0]['lr'] = args.base_lr * 0.1 ** (it // 2500)
        if it % 200 == 0:                                                                 # LA_BCP_train.py:172-186,278-291
            t_v = time.time()
            dice = validate(args, model)
            if dice is not None and dice > best_dice and rank == 0:
                best_dice = round(dice, 4)
                for name in ('iter_{}_dice_{}.pth'.format(it, best_dice), '{}_best_model.pth'.format(args.model)):
                    path = os.path.join(snapshot_path, name)
                    if kind == "la_pre":
                        save_net_opt(model, optimizer, path)
                    else:
                        torch.save(model.state_dict(), path)
                logging.info("save best model to {}".format(os.path.join(snapshot_path, 'iter_{}_dice_{}.pth'.format(it, best_dice))))
            t_io += time.time() - t_v
        if args.ckpt_every and it % args.ckpt_every == 0 and rank == 0:
            t_v = time.time()
            save_resume(resume_path, model, optimizer, ema_model, it, stage, {"best_dice": best_dice})
            t_io += time.time() - t_v
    torch.cuda.synchronize(device)
    if rank == 0:
        logging.info("%s: %d iterations, %.2f it/s" % (stage, it - it0, (it - it0) / max(time.time() - t0 - t_io, 1e-9)))
    return it, best_dice


def pre_train(args, snapshot_path, device, rank):
    from bcp_b200.networks.net_factory import net_factory
    from bcp_b200.optim import FusedSGD_EMA
    model = net_factory(net_type=args.model, in_chns=1, class_num=num_classes, mode="train")
    optimizer = FusedSGD_EMA(model, None, lr=args.base_lr, momentum=0.9, weight_decay=0.0001)
    model.train()
    run_stage(args, "pre_train", model, None, optimizer, snapshot_path, device, rank, args.pre_max_iteration)
    best = os.path.join(snapshot_path, '{}_best_model.pth'.format(args.model))
    if rank == 0 and not os.path.exists(best):           # synthetic runs have no validation to pick a best model
        save_net_opt(model, optimizer, best)
    return model


def self_train(args, pre_snapshot_path, self_snapshot_path, device, rank):
    from bcp_b200.networks.net_factory import net_factory
    from bcp_b200.optim import FusedSGD_EMA
    model = net_factory(net_type=args.model, in_chns=1, class_num=num_classes, mode="train")
    ema_model = net_factory(net_type=args.model, in_chns=1, class_num=num_classes, mode="train")
    for param in ema_model.parameters():
        param.detach_()
    pretrained_model = os.path.join(pre_snapshot_path, f'{args.model}_best_model.pth')
    load_net(model, pretrained_model)
    load_net(ema_model, pretrained_model)
    optimizer = FusedSGD_EMA(model, ema_model, lr=args.base_lr, momentum=0.9, weight_decay=0.0001, ema_alpha=0.99)
    model.train()
    ema_model.train()
    run_stage(args, "self_train", model, ema_model, optimizer, self_snapshot_path, device, rank, args.self_max_iteration)
    best = os.path.join(self_snapshot_path, '{}_best_model.pth'.format(args.model))
    if rank == 0 and not os.path.exists(best):
        torch.save(model.state_dict(), best)


if __name__ == "__main__":
    args = parser.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and "CUDA_VISIBLE_DEVICES" not in os.environ and not args.synthetic:
        os.environ['CUDA_VISIBLE_DEVICES'] = args.gpu
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    if args.deterministic:
        torch.manual_seed(args.seed)
        torch.cuda.manual_seed(args.seed)
        random.seed(args.seed)
        np.random.seed(args.seed + rank)
    pre_snapshot_path = "./model/BCP/LA_{}_{}_labeled/pre_train".format(args.exp, args.labelnum)
    self_snapshot_path = "./model/BCP/LA_{}_{}_labeled/self_train".format(args.exp, args.labelnum)
    if rank == 0:
        for p in (pre_snapshot_path, self_snapshot_path):
            os.makedirs(p, exist_ok=True)
    logging.basicConfig(level=logging.INFO, format='[%(asctime)s.%(msecs)03d] %(message)s', datefmt='%H:%M:%S',
                        handlers=[logging.StreamHandler(sys.stdout)] + ([logging.FileHandler(pre_snapshot_path + "/log.txt")] if rank == 0 else []))
    logging.info(str(args))
    print("Starting BCP training.")
    pre_train(args, pre_snapshot_path, device, rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    self_train(args, pre_snapshot_path, self_snapshot_path, device, rank)
