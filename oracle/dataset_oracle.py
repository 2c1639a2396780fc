"""CPU restatement (numpy) of the reference's LA input pipeline -- TEST INFRASTRUCTURE, never imported by the product.

Follows /root/reference/code/dataloaders/dataset.py: random_rot_flip :52-60, RandomCrop :173-212, RandomRotFlip
:215-225, ToTensor :267-277, TwoStreamBatchSampler + iterate_once / iterate_eternally / grouper :280-307,340-355.
Pinned by tests/golden/dataset.npz, minted by executing the reference's own classes (tests/golden/make_golden.py:
gen_dataset) on seeded volumes; tests/test_oracle_golden.py::test_dataset_pipeline checks this file against it.
"""
import itertools

import numpy as np


def random_rot_flip(image, label):
    k = np.random.randint(0, 4)
    image, label = np.rot90(image, k), np.rot90(label, k)
    axis = np.random.randint(0, 2)
    return np.flip(image, axis=axis).copy(), np.flip(label, axis=axis).copy()


def random_crop(image, label, output_size):
    if label.shape[0] <= output_size[0] or label.shape[1] <= output_size[1] or label.shape[2] <= output_size[2]:
        pw = max((output_size[0] - label.shape[0]) // 2 + 3, 0)
        ph = max((output_size[1] - label.shape[1]) // 2 + 3, 0)
        pd = max((output_size[2] - label.shape[2]) // 2 + 3, 0)
        image = np.pad(image, [(pw, pw), (ph, ph), (pd, pd)], mode="constant", constant_values=0)
        label = np.pad(label, [(pw, pw), (ph, ph), (pd, pd)], mode="constant", constant_values=0)
    w, h, d = image.shape
    w1 = np.random.randint(0, w - output_size[0])
    h1 = np.random.randint(0, h - output_size[1])
    d1 = np.random.randint(0, d - output_size[2])
    sl = (slice(w1, w1 + output_size[0]), slice(h1, h1 + output_size[1]), slice(d1, d1 + output_size[2]))
    return image[sl], label[sl]


def la_train_sample(image, label, output_size):
    """Compose([RandomRotFlip(), RandomCrop(output_size), ToTensor()]) -> (image fp32 [1,X,Y,Z], label int64 [X,Y,Z])."""
    image, label = random_rot_flip(image, label)
    image, label = random_crop(image, label, output_size)
    return image.reshape((1,) + image.shape).astype(np.float32), label.astype(np.int64)


def two_stream_batches(primary, secondary, batch_size, secondary_batch_size):
    """Generator over one epoch of TwoStreamBatchSampler index tuples (primary first)."""
    pb = batch_size - secondary_batch_size
    primary_iter = np.random.permutation(primary)

    def eternal():
        while True:
            yield np.random.permutation(secondary)
    secondary_iter = itertools.chain.from_iterable(eternal())
    a, b = [iter(primary_iter)] * pb, [iter(secondary_iter)] * secondary_batch_size
    for p, s in zip(zip(*a), zip(*b)):
        yield tuple(int(v) for v in p + s)


def la_batch(volumes, indices, output_size):
    imgs, labs = [], []
    for i in indices:
        im, lb = la_train_sample(volumes[i][0], volumes[i][1], output_size)
        imgs.append(im)
        labs.append(lb)
    return np.stack(imgs), np.stack(labs)


def synthetic_la_volumes(n, seed, lo=(20, 18, 14), hi=(40, 36, 30)):
    """Seeded raw 'scans' of varying size (some smaller than the test patch, to exercise the padding branch)."""
    rs = np.random.RandomState(seed)
    vols = []
    for _ in range(n):
        shape = tuple(int(rs.randint(lo[i], hi[i])) for i in range(3))
        vols.append((rs.standard_normal(shape).astype(np.float32), (rs.random_sample(shape) > 0.8).astype(np.uint8)))
    return vols


def synthetic_acdc_slices(n, seed, lo=(40, 36), hi=(72, 70)):
    """Seeded raw 2-D 'slices' of varying size (the ACDC slices are 216x256, 232x256, ...): smooth-ish images in [0, 1]
    and 4-class label maps with connected regions, so nearest-neighbour rotation / zoom moves class boundaries visibly."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        h, w = int(rs.randint(lo[0], hi[0])), int(rs.randint(lo[1], hi[1]))
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        img = (np.sin(yy / rs.uniform(3, 9)) * np.cos(xx / rs.uniform(3, 9)) * 0.5 + 0.5 + 0.05 * rs.standard_normal((h, w))).astype(np.float32)
        cy, cx, r = rs.uniform(0.3, 0.7) * h, rs.uniform(0.3, 0.7) * w, rs.uniform(0.15, 0.3) * min(h, w)
        d = np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
        lab = np.zeros((h, w), np.uint8)
        lab[d < r] = 1
        lab[d < 0.66 * r] = 2
        lab[d < 0.33 * r] = 3
        out.append((img, lab))
    return out
