"""Input pipeline of the LA entry point, API of the reference's ``dataloaders/dataset.py``:
``LAHeart`` (:91-126), ``RandomRotFlip`` (:215-225), ``RandomCrop`` (:173-212), ``ToTensor`` (:267-277) and
``TwoStreamBatchSampler`` (:280-307, with iterate_once / iterate_eternally / grouper :340-355).

The reference reads one h5 file per sample in 4 DataLoader workers, transforms it with numpy on the host and ships
fp32 images + int64 labels (64 MB of labels per LA step) through pinned memory.  Here the whole training set is loaded
ONCE and stays resident in HBM (80 LA volumes are < 2 GB of the 180 GB); a step's batch is cut out of it by one gather
kernel per sample (``bcp_aug_crop_rotflip``) that composes rot90 / flip / zero-pad / crop, with the random draws made
on the host in exactly the reference's ``np.random`` call order (sampler permutations first, then per sample
k, axis, w1, h1, d1).  Labels are uint8 end to end.  No per-step H2D traffic, no worker processes.

Volumes come from ``<base_dir>/2018LA_Seg_Training Set/<name>/mri_norm2.h5`` (needs h5py) or, when h5py is absent,
from ``<base_dir>/<name>.npz`` with the same two arrays ``image`` (fp32 [W,H,D]) and ``label`` (uint8).
"""
from __future__ import annotations

import itertools
import os

import numpy as np
import torch

from .._native import LIB, i3, ptr, stream


# ---- sampler (host logic, identical draws) ----------------------------------------------------------------------------
def iterate_once(iterable):
    return np.random.permutation(iterable)


def iterate_eternally(indices):
    def infinite_shuffles():
        while True:
            yield np.random.permutation(indices)
    return itertools.chain.from_iterable(infinite_shuffles())


def grouper(iterable, n):
    args = [iter(iterable)] * n
    return zip(*args)


class TwoStreamBatchSampler:
    """An 'epoch' is one pass over the primary (labeled) indices; the secondary (unlabeled) ones cycle for ever.
    Batches are primary indices first (dataloaders/dataset.py:280-307)."""

    def __init__(self, primary_indices, secondary_indices, batch_size, secondary_batch_size):
        self.primary_indices = primary_indices
        self.secondary_indices = secondary_indices
        self.secondary_batch_size = secondary_batch_size
        self.primary_batch_size = batch_size - secondary_batch_size
        assert len(self.primary_indices) >= self.primary_batch_size > 0
        assert len(self.secondary_indices) >= self.secondary_batch_size > 0

    def __iter__(self):
        primary_iter = iterate_once(self.primary_indices)
        secondary_iter = iterate_eternally(self.secondary_indices)
        return (primary_batch + secondary_batch
                for (primary_batch, secondary_batch)
                in zip(grouper(primary_iter, self.primary_batch_size), grouper(secondary_iter, self.secondary_batch_size)))

    def __len__(self):
        return len(self.primary_indices) // self.primary_batch_size


# ---- transform parameter draws (host) ---------------------------------------------------------------------------------
def draw_rotflip_crop(shape, output_size, rot_flip=True):
    """The draws RandomRotFlip then RandomCrop make for a volume of ``shape``, in their np.random call order.
    Returns dict(k, axis, pad(3), origin(3))."""
    w, h, d = (int(v) for v in shape)
    k, axis = 0, 0
    flip = False
    if rot_flip:
        k = int(np.random.randint(0, 4))                 # random_rot_flip, dataset.py:53
        axis = int(np.random.randint(0, 2))              # :56
        flip = True
        if k & 1:
            w, h = h, w
    pw = ph = pd = 0
    if w <= output_size[0] or h <= output_size[1] or d <= output_size[2]:          # RandomCrop pads small volumes, :190-199
        pw = max((output_size[0] - w) // 2 + 3, 0)
        ph = max((output_size[1] - h) // 2 + 3, 0)
        pd = max((output_size[2] - d) // 2 + 3, 0)
    w, h, d = w + 2 * pw, h + 2 * ph, d + 2 * pd
    w1 = int(np.random.randint(0, w - output_size[0]))   # :203-205
    h1 = int(np.random.randint(0, h - output_size[1]))
    d1 = int(np.random.randint(0, d - output_size[2]))
    return dict(k=k, axis=axis, flip=flip, pad=(pw, ph, pd), origin=(w1, h1, d1))


class LAHeart:
    """Device-resident LA dataset.  ``split`` / ``num`` as in the reference (:93-111).  ``batch(indices, out_size)`` returns
    {'image': [B,1,X,Y,Z] fp32, 'label': [B,X,Y,Z] uint8} on the device, transformed like
    Compose([RandomRotFlip(), RandomCrop(out_size), ToTensor()]) applied sample by sample in index order."""

    def __init__(self, base_dir=None, split="train", num=None, device=None, volumes=None):
        self._base_dir = base_dir
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        if volumes is not None:                                     # [(image fp32 [W,H,D], label uint8)] already in memory
            self.image_list = ["mem%d" % i for i in range(len(volumes))]
        else:
            path = os.path.join(base_dir, "train.list" if split == "train" else "test.list")
            with open(path, "r") as f:
                self.image_list = [item.replace("\n", "") for item in f.readlines()]
        if num is not None:
            self.image_list = self.image_list[:num]
            if volumes is not None:
                volumes = volumes[:num]
        print("total {} samples".format(len(self.image_list)))
        self.volumes = []
        src = volumes if volumes is not None else (self._read(name) for name in self.image_list)
        for image, label in src:
            img = torch.as_tensor(np.ascontiguousarray(image, dtype=np.float32)).to(self.device)
            lab = torch.as_tensor(np.ascontiguousarray(label).astype(np.uint8)).to(self.device)
            assert img.shape == lab.shape and img.dim() == 3
            self.volumes.append((img, lab))

    def _read(self, name):
        h5 = os.path.join(self._base_dir, "2018LA_Seg_Training Set", name, "mri_norm2.h5")
        if os.path.exists(h5):
            import h5py                      # not part of this image; present wherever the real LA h5 files are
            with h5py.File(h5, "r") as f:
                return f["image"][:], f["label"][:]
        z = np.load(os.path.join(self._base_dir, name + ".npz"))
        return z["image"], z["label"]

    def __len__(self):
        return len(self.image_list)

    def batch(self, indices, output_size, rot_flip=True, out=None):
        B = len(indices)
        ox, oy, oz = (int(v) for v in output_size)
        if out is None:
            out = {"image": torch.empty((B, 1, ox, oy, oz), dtype=torch.float32, device=self.device),
                   "label": torch.empty((B, ox, oy, oz), dtype=torch.uint8, device=self.device)}
        params = []
        for b, idx in enumerate(indices):
            img, lab = self.volumes[int(idx)]
            p = draw_rotflip_crop(img.shape, (ox, oy, oz), rot_flip)
            params.append(p)
            if not p["flip"]:
                # no RandomRotFlip in the pipeline: k = 0 and "flip twice" = identity is not expressible, so un-flip by
                # choosing axis 0 with a mirrored origin is avoided -- the kernel takes flip_axis -1 as "no flip"
                raise NotImplementedError("pipelines without RandomRotFlip are not used by the LA entry point")
            LIB.call("bcp_aug_crop_rotflip", ptr(img), ptr(lab), out["image"][b].data_ptr(), out["label"][b].data_ptr(),
                     i3(*img.shape), i3(ox, oy, oz), p["k"], p["axis"], i3(*p["pad"]), i3(*p["origin"]), stream())
        out["params"] = params
        return out


class TwoStreamLoader:
    """DataLoader(db, batch_sampler=TwoStreamBatchSampler(...)) with num_workers=0 semantics: an iterator over epochs'
    batches; every batch is produced on the device."""

    def __init__(self, db: LAHeart, batch_sampler: TwoStreamBatchSampler, output_size):
        self.db, self.sampler, self.output_size = db, batch_sampler, output_size

    def __iter__(self):
        for indices in self.sampler:
            yield self.db.batch(list(indices), self.output_size)

    def __len__(self):
        return len(self.sampler)


# ======================================================================================================================
# ACDC (2-D) input pipeline: ``BaseDataSets`` (dataloaders/dataset.py:15-50), ``random_rot_flip`` (:52-59),
# ``random_rotate`` (:62-66), ``RandomGenerator`` (:69-88).
#
# The slices have different sizes (216x256, 232x256, ...) and the reference resamples every one of them to the patch size
# with scipy's nearest-neighbour ``zoom`` / ``rotate`` in DataLoader workers.  This mirror keeps that arithmetic on the host
# -- the SAME scipy calls, so batches are bit-identical to the reference's under the same ``random`` / ``np.random`` streams
# (num_workers=0 order; tests/golden/acdc_dataset.npz) -- but reads the training set once into host memory, emits uint8
# labels, and hands batches over in pinned double buffers that the step's H2D copy stream consumes (bcp_b200/graph.py).
# A device-side resampler is the next step for this row; at 24 slices of 256x256 per step the host
# transform is ~2 ms of numpy per batch, so ``SliceLoader`` prepares batch k+1 on a worker thread while step k runs.
# ======================================================================================================================
def random_rot_flip(image, label):
    """rot90 by a random quarter turn, then a flip along a random axis (draws: randint(0,4), randint(0,2))."""
    k = np.random.randint(0, 4)
    axis = np.random.randint(0, 2)          # drawn after the rotation in the reference, which consumes no randomness in between
    image, label = np.rot90(image, k), np.rot90(label, k)
    return np.flip(image, axis=axis).copy(), np.flip(label, axis=axis).copy()


def random_rotate(image, label):
    """nearest-neighbour rotation by an integer angle in [-20, 20), output shape kept."""
    from scipy import ndimage
    angle = np.random.randint(-20, 20)
    return (ndimage.rotate(image, angle, order=0, reshape=False),
            ndimage.rotate(label, angle, order=0, reshape=False))


class RandomGenerator:
    """Augment one {'image','label'} slice and resample it to ``output_size`` (nearest neighbour).  Draw order as in the
    reference: ``random.random()`` > 0.5 -> rot/flip, else a second ``random.random()`` > 0.5 -> small rotation.
    ``draw()`` makes exactly those draws and ``apply()`` is deterministic given them, so a loader can draw in the reference's
    order on one thread and run the (expensive) scipy resampling of a batch in parallel workers (``SliceLoader(workers=)``)."""

    def __init__(self, output_size):
        self.output_size = output_size

    @staticmethod
    def draw():
        import random
        if random.random() > 0.5:
            k = int(np.random.randint(0, 4))
            return ("rot_flip", k, int(np.random.randint(0, 2)))
        if random.random() > 0.5:
            return ("rotate", int(np.random.randint(-20, 20)))
        return ("none",)

    def apply(self, image, label, params):
        from scipy import ndimage
        from scipy.ndimage import zoom
        if params[0] == "rot_flip":
            _, k, axis = params
            image, label = np.flip(np.rot90(image, k), axis=axis).copy(), np.flip(np.rot90(label, k), axis=axis).copy()
        elif params[0] == "rotate":
            image = ndimage.rotate(image, params[1], order=0, reshape=False)
            label = ndimage.rotate(label, params[1], order=0, reshape=False)
        x, y = image.shape
        fx, fy = self.output_size[0] / x, self.output_size[1] / y
        return zoom(image, (fx, fy), order=0).astype(np.float32), zoom(label, (fx, fy), order=0).astype(np.uint8)

    def __call__(self, sample):
        image, label = self.apply(sample["image"], sample["label"], self.draw())
        return {"image": torch.from_numpy(image).unsqueeze(0), "label": torch.from_numpy(label)}


class BaseDataSets:
    """ACDC slices ('train': ``train_slices.list`` -> ``data/slices/<case>.h5``) or volumes ('val': ``val.list`` ->
    ``data/<case>.h5``), read ONCE into host memory (h5py where available, else ``<case>.npz`` next to where the h5 would
    be).  ``slices=[(image, label), ...]`` serves in-memory data (tests, synthetic runs)."""

    def __init__(self, base_dir=None, split="train", num=None, transform=None, slices=None):
        self._base_dir, self.split, self.transform = base_dir, split, transform
        if slices is not None:
            self.sample_list = ["mem%d" % i for i in range(len(slices))]
        else:
            name = "train_slices.list" if split == "train" else "val.list"
            with open(os.path.join(base_dir, name), "r") as f:
                self.sample_list = [item.replace("\n", "") for item in f.readlines()]
        if num is not None and split == "train":
            self.sample_list = self.sample_list[:num]
            if slices is not None:
                slices = slices[:num]
        print("total {} samples".format(len(self.sample_list)))
        self._data = list(slices) if slices is not None else [self._read(c) for c in self.sample_list]

    def _read(self, case):
        stem = os.path.join(self._base_dir, "data", "slices" if self.split == "train" else "", case)
        if os.path.exists(stem + ".h5"):
            import h5py                      # not part of this image; present wherever the real ACDC h5 files are
            with h5py.File(stem + ".h5", "r") as f:
                return f["image"][:], f["label"][:]
        z = np.load(stem + ".npz")
        return z["image"], z["label"]

    def __len__(self):
        return len(self.sample_list)

    def __getitem__(self, idx):
        image, label = self._data[idx]
        sample = {"image": image, "label": label}
        if self.split == "train" and self.transform is not None:
            sample = self.transform(sample)
        sample["case"] = self.sample_list[idx]
        return sample


def patients_to_slices(dataset, patiens_num):
    """Labeled-slice count for a number of labeled patients (ACDC_BCP_train.py:70-79)."""
    if "ACDC" in dataset:
        ref_dict = {"1": 32, "3": 68, "7": 136, "14": 256, "21": 396, "28": 512, "35": 664, "140": 1312}
    elif "Prostate" in dataset:
        ref_dict = {"2": 27, "4": 53, "8": 120, "12": 179, "16": 256, "21": 312, "42": 623}
    else:
        raise ValueError("unknown dataset %r" % (dataset,))
    return ref_dict[str(patiens_num)]


_WORKER = {}


def _worker_init(data, output_size, sh_img, sh_lab):
    _WORKER["data"], _WORKER["gen"] = data, RandomGenerator(output_size)
    _WORKER["img"], _WORKER["lab"] = sh_img.numpy(), sh_lab.numpy()          # shared-memory batch, inherited through fork


def _worker_apply(job):
    b, idx, params = job
    image, label = _WORKER["data"][idx]
    im, lb = _WORKER["gen"].apply(image, label, params)
    _WORKER["img"][b, 0] = im                                                 # results go straight into the shared batch:
    _WORKER["lab"][b] = lb                                                    # nothing but the job tuple crosses the pipe
    return b


class SliceLoader:
    """``DataLoader(db, batch_sampler=TwoStreamBatchSampler(...), num_workers=0)`` for the 2-D pipeline: yields
    {'image': [B,1,H,W] fp32, 'label': [B,H,W] uint8} in PINNED host memory so the graphed step's copy stream can take it
    without a staging copy.  The buffers form a ring of RING = 4: ``GraphedStep.load`` of batch k+2 returns only after the
    H2D copy of batch k has finished (bcp_b200/graph.py), and with ``prefetch=True`` batch k+4 is being written while the
    consumer is at most inside ``load`` of batch k+2 -- so a buffer is never rewritten under a copy in flight.
    ``prefetch=True`` builds the next batch on a worker thread; the draws stay in the single-process order because only that
    thread touches the RNG streams while it runs.  ``workers=N`` (the dataset's transform must be a ``RandomGenerator``): the
    random draws of a batch are still made here, sample by sample in the reference's order, but the scipy resampling (which
    holds the GIL) runs in N forked worker processes that hold the raw slices and write into one shared-memory batch -- same
    batches, bit for bit."""
    RING = 4

    def __init__(self, db: BaseDataSets, batch_sampler: TwoStreamBatchSampler, pin=None, prefetch=False, workers=0):
        self.db, self.sampler, self.prefetch = db, batch_sampler, prefetch
        self.pin = torch.cuda.is_available() if pin is None else pin
        self._bufs, self._turn = None, 0
        self._pool = None
        if workers > 0:
            assert isinstance(db.transform, RandomGenerator) and db.split == "train", "workers need a RandomGenerator transform"
            import multiprocessing as mp
            B = batch_sampler.primary_batch_size + batch_sampler.secondary_batch_size
            H, W = (int(v) for v in db.transform.output_size)
            self._sh_img = torch.empty((B, 1, H, W), dtype=torch.float32).share_memory_()
            self._sh_lab = torch.empty((B, H, W), dtype=torch.uint8).share_memory_()
            self._pool = mp.get_context("fork").Pool(workers, initializer=_worker_init,
                                                     initargs=(db._data, db.transform.output_size, self._sh_img, self._sh_lab))

    def close(self):
        if self._pool is not None:
            self._pool.terminate()
            self._pool = None

    def __del__(self):
        self.close()

    def _next_buffers(self, img_shape, lab_shape):
        if self._bufs is None:
            self._bufs = [(torch.empty(tuple(img_shape), dtype=torch.float32, pin_memory=self.pin),
                           torch.empty(tuple(lab_shape), dtype=torch.uint8, pin_memory=self.pin)) for _ in range(self.RING)]
        img, lab = self._bufs[self._turn]
        self._turn = (self._turn + 1) % self.RING
        return img, lab

    def _collate(self, indices):
        if self._pool is not None:
            jobs = [(b, int(i), RandomGenerator.draw()) for b, i in enumerate(indices)]      # draws in index order, this thread only
            self._pool.map(_worker_apply, jobs)
            img, lab = self._next_buffers(self._sh_img.shape, self._sh_lab.shape)
            img.copy_(self._sh_img)
            lab.copy_(self._sh_lab)
            return {"image": img, "label": lab, "case": [self.db.sample_list[i] for _, i, _ in jobs]}
        samples = [self.db[int(i)] for i in indices]
        img0, lab0 = samples[0]["image"], samples[0]["label"]
        img, lab = self._next_buffers((len(samples),) + tuple(img0.shape), (len(samples),) + tuple(lab0.shape))
        for b, s in enumerate(samples):
            img[b].copy_(s["image"])
            lab[b].copy_(s["label"])
        return {"image": img, "label": lab, "case": [s["case"] for s in samples]}

    def __iter__(self):
        it = iter(self.sampler)
        if not self.prefetch:
            for indices in it:
                yield self._collate(indices)
            return
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1) as pool:
            nxt = next(it, None)
            fut = pool.submit(self._collate, nxt) if nxt is not None else None
            while fut is not None:
                batch = fut.result()
                nxt = next(it, None)                    # the sampler's draws stay between two batches' transform draws
                fut = pool.submit(self._collate, nxt) if nxt is not None else None
                yield batch

    def __len__(self):
        return len(self.sampler)
