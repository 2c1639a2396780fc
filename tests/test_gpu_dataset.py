"""Device-side input pipeline: bcp_b200.dataloaders.dataset against the batches the reference's
own LAHeart / RandomRotFlip / RandomCrop / ToTensor / TwoStreamBatchSampler produced under a single-process DataLoader
(tests/golden/dataset.npz, minted by tests/golden/make_golden.py: gen_dataset).  Pure data movement: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import dataset_oracle as D
from tests.util import load_golden

pytestmark = pytest.mark.gpu


def test_la_pipeline_bit_exact():
    from bcp_b200.dataloaders.dataset import LAHeart, TwoStreamBatchSampler, TwoStreamLoader
    g = load_golden("dataset")
    vols = D.synthetic_la_volumes(int(g["nvol"]), 4242)
    patch = tuple(int(v) for v in g["patch"])
    lab_n, bs, lbs = int(g["labeled"]), int(g["batch_size"]), int(g["labeled_bs"])
    db = LAHeart(volumes=vols, device=torch.device("cuda:0"))
    loader = TwoStreamLoader(db, TwoStreamBatchSampler(list(range(lab_n)), list(range(lab_n, len(vols))), bs, bs - lbs), patch)
    assert len(loader) == lab_n // lbs
    np.random.seed(int(g["seed"]))
    b = 0
    for _ in range(int(g["epochs"])):
        for batch in loader:
            assert batch["image"].dtype == torch.float32 and batch["label"].dtype == torch.uint8
            assert batch["image"].shape == (bs, 1) + patch and batch["label"].shape == (bs,) + patch
            assert np.array_equal(batch["image"].cpu().numpy(), g[f"b{b}_image"]), b
            assert np.array_equal(batch["label"].cpu().numpy(), g[f"b{b}_label"]), b
            b += 1
    assert b == int(g["nbatches"])


@pytest.mark.parametrize("k,axis", [(0, 0), (1, 1), (2, 0), (3, 1), (1, 0), (3, 0)])
def test_crop_rotflip_kernel_all_orientations(k, axis):
    """Every (rot90 k, flip axis) pair on a non-cubic volume, with and without the zero padding of small volumes, at an LA-size
    patch, against numpy."""
    from bcp_b200._native import LIB, i3, ptr, stream
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(10 * k + axis)
    for shape, patch in (((150, 131, 88), (112, 112, 80)), ((100, 120, 70), (112, 112, 80))):
        im = rs.standard_normal(shape).astype(np.float32)
        lb = (rs.random_sample(shape) > 0.7).astype(np.uint8)
        r_im, r_lb = np.flip(np.rot90(im, k), axis=axis), np.flip(np.rot90(lb, k), axis=axis)
        pad = (0, 0, 0)
        if any(r_im.shape[i] <= patch[i] for i in range(3)):
            pad = tuple(max((patch[i] - r_im.shape[i]) // 2 + 3, 0) for i in range(3))
        r_im = np.pad(r_im, [(p, p) for p in pad], mode="constant")
        r_lb = np.pad(r_lb, [(p, p) for p in pad], mode="constant")
        org = tuple(int(rs.randint(0, r_im.shape[i] - patch[i])) for i in range(3))
        sl = tuple(slice(org[i], org[i] + patch[i]) for i in range(3))
        t_im, t_lb = torch.from_numpy(im).to(dev), torch.from_numpy(lb).to(dev)
        o_im = torch.empty(patch, dtype=torch.float32, device=dev)
        o_lb = torch.empty(patch, dtype=torch.uint8, device=dev)
        LIB.call("bcp_aug_crop_rotflip", ptr(t_im), ptr(t_lb), ptr(o_im), ptr(o_lb), i3(*shape), i3(*patch), k, axis, i3(*pad),
                 i3(*org), stream())
        assert np.array_equal(o_im.cpu().numpy(), r_im[sl]) and np.array_equal(o_lb.cpu().numpy(), r_lb[sl])
