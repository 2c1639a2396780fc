#!/usr/bin/env python
"""LA (left atrium) BCP training entry point on the B200-native kernels.

Keeps the CLI flags, defaults, snapshot directory layout and the two-stage schedule of the reference's
``code/LA_BCP_train.py`` (/root/reference/code/LA_BCP_train.py:32-55,351-371): stage 1 ``pre_train`` (copy-paste
between two labeled volumes), stage 2 ``self_train`` (EMA teacher pseudo-labels, bidirectional copy-paste, SGD, EMA).
The step bodies live in ``bcp_b200.step``.  Extra flags: ``--synthetic`` (seeded synthetic volumes of the LA shape;
the real LA h5 data and h5py are not available in this environment), ``--max_steps`` (bound both stages),
data-parallel launch via ``torchrun`` (one process per GPU, one NCCL all-reduce of the flat gradient per step).

Out of scope here: tensorboard images, the sliding-window validation every 200 iterations (needs the
real data and medpy) -- validation is skipped in synthetic mode.
"""
import argparse
import logging
import os
import random
import sys

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

patch_size = (112, 112, 80)
num_classes = 2


class SyntheticLA:
    """Seeded stand-in for LAHeart + RandomRotFlip/RandomCrop/ToTensor + TwoStreamBatchSampler
    (dataloaders/dataset.py:91-126,280-307): yields {'image': [B,1,112,112,80] fp32, 'label': [B,112,112,80] uint8},
    labeled samples first."""

    def __init__(self, batch_size, seed, device):
        self.bs, self.dev = batch_size, device
        self.gen = torch.Generator(device="cpu").manual_seed(seed)

    def __iter__(self):
        while True:
            img = torch.randn((self.bs, 1) + patch_size, generator=self.gen)
            noise = torch.randn((self.bs, 1) + patch_size, generator=self.gen)
            sm = torch.nn.functional.avg_pool3d(noise, 5, stride=1, padding=2)[:, 0]
            lab = (sm > sm.std()).to(torch.uint8)
            yield {"image": img.pin_memory().to(self.dev, non_blocking=True), "label": lab.pin_memory().to(self.dev, non_blocking=True)}


def make_loader(args, device, rank):
    if args.synthetic:
        return SyntheticLA(args.batch_size, args.seed + rank, device)
    raise RuntimeError("real LA data needs h5py and the dataset at --root_path; neither exists in this environment "
                       "(use --synthetic 1)")


def save_net_opt(net, optimizer, path):
    torch.save({'net': net.state_dict(), 'opt': optimizer.state_dict()}, str(path))      # LA_BCP_train.py:79-84


def load_net(net, path):
    net.load_state_dict(torch.load(str(path))['net'])                                    # LA_BCP_train.py:91-93


def pre_train(args, snapshot_path, device, rank):
    from bcp_b200.networks.net_factory import net_factory
    from bcp_b200.optim import FusedSGD_EMA
    from bcp_b200.step import la_pre_train_step
    model = net_factory(net_type=args.model, in_chns=1, class_num=num_classes, mode="train")
    optimizer = FusedSGD_EMA(model, None, lr=args.base_lr, momentum=0.9, weight_decay=0.0001)
    model.train()
    iters = args.pre_max_iteration if not args.max_steps else min(args.max_steps, args.pre_max_iteration)
    it = 0
    for batch in make_loader(args, device, rank):
        r = la_pre_train_step(model, optimizer, batch['image'], batch['label'], args.labeled_bs, args.mask_ratio)
        it += 1
        if it % args.log_every == 0 and rank == 0:
            logging.info('iteration %d : loss: %03f, loss_dice: %03f, loss_ce: %03f' % (it, float(r['loss']), float(r['loss_dice']), float(r['loss_ce'])))
        if it >= iters:
            break
    if rank == 0:
        save_net_opt(model, optimizer, os.path.join(snapshot_path, '{}_best_model.pth'.format(args.model)))
    return model


def self_train(args, pre_snapshot_path, self_snapshot_path, device, rank):
    from bcp_b200.networks.net_factory import net_factory
    from bcp_b200.optim import FusedSGD_EMA
    from bcp_b200.step import la_self_train_step
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
    iters = args.self_max_iteration if not args.max_steps else min(args.max_steps, args.self_max_iteration)
    it = 0
    for batch in make_loader(args, device, rank):
        r = la_self_train_step(model, ema_model, optimizer, batch['image'], batch['label'], args.labeled_bs, args.mask_ratio, args.u_weight)
        it += 1
        if it % args.log_every == 0 and rank == 0:
            logging.info('iteration %d : loss: %03f, loss_l: %03f, loss_u: %03f' % (it, float(r['loss']), float(r['loss_l']), float(r['loss_u'])))
        if it % 2500 == 0:                                                            # LA_BCP_train.py:273-276
            optimizer.param_groups[0]['lr'] = args.base_lr * 0.1 ** (it // 2500)
        if it >= iters:
            break
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(self_snapshot_path, '{}_best_model.pth'.format(args.model)))


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
