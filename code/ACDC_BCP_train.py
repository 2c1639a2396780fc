#!/usr/bin/env python
"""ACDC BCP training entry point on the B200-native kernels.

Keeps the CLI flags/defaults and the two-stage schedule of the reference's ``code/ACDC_BCP_train.py``
(/root/reference/code/ACDC_BCP_train.py:33-56,445-477).  The self-training step body is
``bcp_b200.step.acdc_self_train_step`` (reference :354-390); pre-training (:237-255) mixes two labeled slices and reuses
``mix_loss(u_weight=1.0, unlab=True)`` exactly like the reference.  ``--synthetic 1`` (default) generates seeded 256x256
slices because the ACDC data is not available here; ``--synthetic 0`` reads ``--root_path`` through
``bcp_b200.dataloaders.dataset`` (BaseDataSets / RandomGenerator / TwoStreamBatchSampler, bit-exact against the reference's
batches: tests/golden/acdc_dataset.npz).
"""
import argparse
import logging
import os
import random
import time
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

parser = argparse.ArgumentParser()
parser.add_argument('--root_path', type=str, default='/data/byh_data/SSNet_data/ACDC', help='Name of Experiment')
parser.add_argument('--exp', type=str, default='BCP', help='experiment_name')
parser.add_argument('--model', type=str, default='unet', help='model_name')
parser.add_argument('--pre_iterations', type=int, default=10000, help='maximum epoch number to train')
parser.add_argument('--max_iterations', type=int, default=30000, help='maximum epoch number to train')
parser.add_argument('--batch_size', type=int, default=24, help='batch_size per gpu')
parser.add_argument('--deterministic', type=int, default=1, help='whether use deterministic training')
parser.add_argument('--base_lr', type=float, default=0.01, help='segmentation network learning rate')
parser.add_argument('--patch_size', type=list, default=[256, 256], help='patch size of network input')
parser.add_argument('--seed', type=int, default=1337, help='random seed')
parser.add_argument('--num_classes', type=int, default=4, help='output channel of network')
parser.add_argument('--labeled_bs', type=int, default=12, help='labeled_batch_size per gpu')
parser.add_argument('--labelnum', type=int, default=7, help='labeled data')
parser.add_argument('--u_weight', type=float, default=0.5, help='weight of unlabeled pixels')
parser.add_argument('--gpu', type=str, default='0', help='GPU to use')
parser.add_argument('--consistency', type=float, default=0.1, help='consistency')
parser.add_argument('--consistency_rampup', type=float, default=200.0, help='consistency_rampup')
parser.add_argument('--magnitude', type=float, default=6.0, help='magnitude')
parser.add_argument('--s_param', type=int, default=6, help='multinum of random masks')
parser.add_argument('--synthetic', type=int, default=1)
parser.add_argument('--loader_workers', type=int, default=0,
                    help='--synthetic 0: processes for the host resampling (scipy rotate/zoom); 0 = on the prefetch thread')
parser.add_argument('--max_steps', type=int, default=0)
parser.add_argument('--log_every', type=int, default=10)
parser.add_argument('--graph', type=int, default=1, help='replay each step as one CUDA graph (0: eager step functions)')
parser.add_argument('--resume', type=int, default=0, help='continue a stage from <snapshot>/resume.pth when it exists')
parser.add_argument('--ckpt_every', type=int, default=0, help='write the resume artefact every N iterations (0: never)')


class SyntheticACDC:
    """{'image': [B,1,256,256] fp32 in [0,1], 'label': [B,256,256] uint8 in 0..3}, labeled slices first
    (dataloaders/dataset.py:15-50,69-88,280-307).  A few pinned host batches are cycled: the per-step H2D copy stays, the
    host RNG does not sit in the training loop."""

    def __init__(self, batch_size, patch, seed, device, pool=4):
        self.bs, self.patch, self.dev = batch_size, tuple(patch), device
        gen = torch.Generator(device="cpu").manual_seed(seed)
        self.pool = []
        for _ in range(pool):
            img = torch.rand((self.bs, 1) + self.patch, generator=gen)
            noise = torch.randn((self.bs, 1) + self.patch, generator=gen)
            sm = torch.nn.functional.avg_pool2d(noise, 9, stride=1, padding=4)[:, 0]
            sm = sm / sm.std()
            lab = torch.bucketize(sm, torch.tensor([0.6, 1.0, 1.5])).to(torch.uint8)
            self.pool.append((img.pin_memory(), lab.pin_memory()))

    def __iter__(self):
        k = 0
        while True:
            img, lab = self.pool[k % len(self.pool)]
            k += 1
            yield {"image": img, "label": lab}


def make_loader(args, device, rank):
    """--synthetic 0: the reference's loader construction (ACDC_BCP_train.py:207-219 / :318-330) on
    bcp_b200.dataloaders.dataset -- slices read once into host memory (h5py, or <case>.npz where h5py is absent), the
    reference's RandomGenerator arithmetic on a prefetch thread, uint8 labels, pinned ring buffers for the step's H2D
    copy stream; epochs are chained for ever (the stage loop stops at its iteration budget)."""
    if args.synthetic:
        return SyntheticACDC(args.batch_size, args.patch_size, args.seed + rank, device)
    from bcp_b200.dataloaders.dataset import BaseDataSets, RandomGenerator, SliceLoader, TwoStreamBatchSampler, patients_to_slices
    db_train = BaseDataSets(base_dir=args.root_path, split="train", num=None, transform=RandomGenerator(args.patch_size))
    total_slices = len(db_train)
    labeled_slice = patients_to_slices(args.root_path, args.labelnum)
    logging.info("Total slices is: {}, labeled slices is:{}".format(total_slices, labeled_slice))
    batch_sampler = TwoStreamBatchSampler(list(range(0, labeled_slice)), list(range(labeled_slice, total_slices)), args.batch_size,
                                          args.batch_size - args.labeled_bs)
    loader = SliceLoader(db_train, batch_sampler, prefetch=True, workers=args.loader_workers)

    def epochs():
        while True:
            for batch in loader:
                yield batch
    return epochs()


def run_stage(args, stage, model, ema_model, optimizer, snapshot_path, device, rank, max_iterations):
    """One stage's loop.  Every step is ONE CUDA-graph replay (bcp_b200/graph.py): the U-Net step is ~330 launches of a few
    microseconds each, so the eager step functions (--graph 0) are bound by the Python enqueue, not by the GPU."""
    from bcp_b200.graph import GraphedStep
    from bcp_b200.step import acdc_pre_train_step, acdc_self_train_step
    from bcp_b200.utils.checkpoint import load_resume, save_resume
    kind = "acdc_pre" if stage == "pre_train" else "acdc"
    iters = max_iterations if not args.max_steps else min(args.max_steps, max_iterations)
    resume_path = os.path.join(snapshot_path, "resume.pth")
    it = 0
    if args.resume and os.path.exists(resume_path):
        it, _, _ = load_resume(resume_path, model, optimizer, ema_model)
        logging.info("resumed %s at iteration %d from %s" % (stage, it, resume_path))
    gs = None
    if args.graph:
        kw = dict(labeled_bs=args.labeled_bs)
        if kind == "acdc":
            kw["u_weight"] = args.u_weight
        state = np.random.get_state()                  # the capture warm-up draws boxes
        gs = GraphedStep(kind, model, ema_model, optimizer, (args.batch_size, 1) + tuple(args.patch_size), device=device, **kw)
        np.random.set_state(state)
    loader = make_loader(args, device, rank)
    torch.cuda.synchronize(device)
    t0, it0, t_io = time.time(), it, 0.0          # t_io: validation + checkpoint time, excluded from the reported step rate
    for batch in loader:
        if it >= iters:
            break
        if gs is not None:
            r = gs(batch['image'], batch['label'])
        else:
            vol, lab = batch['image'].to(device, non_blocking=True), batch['label'].to(device, non_blocking=True)
            if kind == "acdc_pre":
                r = acdc_pre_train_step(model, optimizer, vol, lab, args.labeled_bs)           # ACDC_BCP_train.py:237-255
            else:
                r = acdc_self_train_step(model, ema_model, optimizer, vol, lab, args.labeled_bs, args.u_weight)   # :354-390
        it += 1
        if it % args.log_every == 0 and rank == 0:
            logging.info('iteration %d: loss: %f, mix_dice: %f, mix_ce: %f' % (it, float(r['loss']), float(r['loss_dice']), float(r['loss_ce'])))
        if args.ckpt_every and it % args.ckpt_every == 0 and rank == 0:
            t_v = time.time()
            save_resume(resume_path, model, optimizer, ema_model, it, stage)
            t_io += time.time() - t_v
    torch.cuda.synchronize(device)
    if rank == 0:
        logging.info("%s: %d iterations, %.2f it/s" % (stage, it - it0, (it - it0) / max(time.time() - t0 - t_io, 1e-9)))
    return it


def pre_train(args, snapshot_path, device, rank):
    from bcp_b200.networks.net_factory import BCP_net
    from bcp_b200.optim import FusedSGD_EMA
    model = BCP_net(in_chns=1, class_num=args.num_classes)
    optimizer = FusedSGD_EMA(model, None, lr=args.base_lr, momentum=0.9, weight_decay=0.0001)
    model.train()
    run_stage(args, "pre_train", model, None, optimizer, snapshot_path, device, rank, args.pre_iterations)
    if rank == 0:
        torch.save({'net': model.state_dict(), 'opt': optimizer.state_dict()}, os.path.join(snapshot_path, '{}_best_model.pth'.format(args.model)))


def self_train(args, pre_snapshot_path, snapshot_path, device, rank):
    from bcp_b200.networks.net_factory import BCP_net
    from bcp_b200.optim import FusedSGD_EMA
    model = BCP_net(in_chns=1, class_num=args.num_classes)
    ema_model = BCP_net(in_chns=1, class_num=args.num_classes, ema=True)
    state = torch.load(os.path.join(pre_snapshot_path, '{}_best_model.pth'.format(args.model)), weights_only=False)
    model.load_state_dict(state['net'])
    ema_model.load_state_dict(state['net'])
    optimizer = FusedSGD_EMA(model, ema_model, lr=args.base_lr, momentum=0.9, weight_decay=0.0001, ema_alpha=0.99, ema_mode="state_dict")
    optimizer.load_state_dict(state['opt'])                                          # load_net_opt(model, optimizer, ...) :335
    model.train()
    ema_model.train()
    run_stage(args, "self_train", model, ema_model, optimizer, snapshot_path, device, rank, args.max_iterations)
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(snapshot_path, '{}_best_model.pth'.format(args.model)))


if __name__ == "__main__":
    args = parser.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    if args.deterministic:
        random.seed(args.seed + rank)            # rank 0 = the reference's streams; replicas draw different batches
        np.random.seed(args.seed + rank)
        torch.manual_seed(args.seed)
        torch.cuda.manual_seed(args.seed)
    pre_snapshot_path = "./model/BCP/ACDC_{}_{}_labeled/pre_train".format(args.exp, args.labelnum)
    self_snapshot_path = "./model/BCP/ACDC_{}_{}_labeled/self_train".format(args.exp, args.labelnum)
    if rank == 0:
        for p in (pre_snapshot_path, self_snapshot_path):
            os.makedirs(p, exist_ok=True)
    logging.basicConfig(level=logging.INFO, format='[%(asctime)s.%(msecs)03d] %(message)s', datefmt='%H:%M:%S',
                        handlers=[logging.StreamHandler(sys.stdout)])
    logging.info(str(args))
    pre_train(args, pre_snapshot_path, device, rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    self_train(args, pre_snapshot_path, self_snapshot_path, device, rank)
