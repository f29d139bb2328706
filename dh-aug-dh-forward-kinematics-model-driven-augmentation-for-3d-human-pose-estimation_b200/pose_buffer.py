"""Device-resident fake-pair bank.

The reference appends every iteration's ``pos_3d_cam`` / ``uv`` / ``cam_para_temp`` to host lists
(models_Fk_GAN/model_fk_gan_train.py:486-488), concatenates them at the end of the GAN epoch and re-serves them
to ``train_posenet`` through ``DataLoader(PoseDataSet(...), batch_size, shuffle=True, pin_memory=True)``
(:504-510; common/data_loader.py:9-36) -- one D2H copy per iteration and one H2D copy per batch.

``DevicePoseBuffer`` keeps the pairs in HBM (bounded ring, sized for 180 GB per GPU: 356 B per pose) and
``DevicePoseBuffer.loader()`` yields batches with the same wire format and -- for the same torch seed -- the SAME
sample order as the reference's shuffled DataLoader under the installed torch: the iterator draws a base seed,
RandomSampler draws its own int64 seed from torch's default RNG and permutes with ``torch.randperm(n, generator=g)``
(torch/utils/data/{dataloader,sampler}.py), which ``shuffled_order`` reproduces draw for draw.  A mini-batch is one gather launch (dhfk_bank_gather).
"""
from __future__ import annotations

import torch

from . import _cabi
from .functional import _require_cuda, _stream_ptr


def shuffled_order(n: int) -> torch.Tensor:
    """The permutation a `DataLoader(dataset_of_n, shuffle=True, num_workers=0)` walks in its next epoch.  Consumes
    torch's default CPU RNG exactly as torch >= 1.x does when such a loader is iterated: one int64 `random_()` draw
    for the iterator's base seed (_BaseDataLoaderIter.__init__), then one for RandomSampler's private generator."""
    torch.empty((), dtype=torch.int64).random_()                      # DataLoader iterator base seed (unused here)
    seed = int(torch.empty((), dtype=torch.int64).random_().item())   # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


def gather_pairs(bank3d, bank2d, bank_cam, idx, rows=None):
    """(bank3d[idx], bank2d[idx], bank_cam[idx]) in one launch; idx int64 on the device; `rows` = valid bank rows."""
    _require_cuda()
    lib = _cabi.load()
    device = bank3d.device
    idx = idx.to(device=device, dtype=torch.int64).contiguous()
    nb = idx.shape[0]
    rows = bank3d.shape[0] if rows is None else int(rows)
    o3 = torch.empty((nb, 16, 3), dtype=torch.float32, device=device)
    o2 = torch.empty((nb, 16, 2), dtype=torch.float32, device=device)
    oc = torch.empty((nb, bank_cam.shape[1]), dtype=torch.float32, device=device) if bank_cam is not None else None
    for t in (bank3d, bank2d) + ((bank_cam,) if bank_cam is not None else ()):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("bank tensors must be contiguous float32 CUDA tensors")
    with torch.cuda.device(device):
        rc = lib.dhfk_bank_gather(bank3d.data_ptr(), bank2d.data_ptr(),
                                  bank_cam.data_ptr() if bank_cam is not None else None,
                                  bank_cam.shape[1] if bank_cam is not None else 0, idx.data_ptr(), nb, rows,
                                  o3.data_ptr(), o2.data_ptr(), oc.data_ptr() if oc is not None else None,
                                  _stream_ptr(device))
    _cabi.check(rc, "dhfk_bank_gather")
    return o3, o2, oc


class DevicePoseBuffer:
    """Append-only (ring when full) bank of (pose3d_cam [16,3], pose2d [16,2], cam [cam_cols]) rows in HBM."""

    def __init__(self, capacity: int, device=None, cam_cols: int = 9):
        _require_cuda()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.capacity, self.cam_cols = int(capacity), int(cam_cols)
        self.pose3d = torch.empty((self.capacity, 16, 3), dtype=torch.float32, device=self.device)
        self.pose2d = torch.empty((self.capacity, 16, 2), dtype=torch.float32, device=self.device)
        self.cam = torch.empty((self.capacity, self.cam_cols), dtype=torch.float32, device=self.device)
        self.reset()

    def reset(self):
        self.count = 0       # rows ever appended since reset
        self._head = 0       # next row to write

    def __len__(self):
        return min(self.count, self.capacity)

    def append(self, pose3d, pose2d, cam):
        """Stores a detached copy of one iteration's outputs (what the reference's `.detach().cpu().numpy()` keeps);
        device-to-device, asynchronous on the current stream.  When the ring is full the oldest rows are replaced."""
        n = pose3d.shape[0]
        p3 = pose3d.detach().reshape(n, 16, 3)
        p2 = pose2d.detach().reshape(n, 16, 2)
        cm = cam.detach().reshape(n, -1)[:, :self.cam_cols]
        if n > self.capacity:
            p3, p2, cm, n = p3[-self.capacity:], p2[-self.capacity:], cm[-self.capacity:], self.capacity
        first = min(n, self.capacity - self._head)
        for dst, src in ((self.pose3d, p3), (self.pose2d, p2), (self.cam, cm)):
            dst[self._head:self._head + first].copy_(src[:first], non_blocking=True)
            if first < n:
                dst[:n - first].copy_(src[first:], non_blocking=True)
        self._head = (self._head + n) % self.capacity
        self.count += n

    def loader(self, batch_size: int, shuffle: bool = True, drop_last: bool = False):
        return DeviceBatchLoader(self, int(batch_size), shuffle, drop_last)


class DeviceBatchLoader:
    """Iterable with the DataLoader surface `train_posenet` uses (`len()`, iteration yielding
    (targets_3d, inputs_2d, action, cam_param), function_aug/model_pos_train.py:24-34); tensors are already on the
    device, so the consumer's `.to(device)` is a no-op."""

    def __init__(self, bank: DevicePoseBuffer, batch_size: int, shuffle: bool, drop_last: bool):
        self.bank, self.batch_size, self.shuffle, self.drop_last = bank, batch_size, shuffle, drop_last

    def __len__(self):
        n = len(self.bank)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.bank)
        order = shuffled_order(n) if self.shuffle else torch.arange(n)
        order = order.to(self.bank.device, non_blocking=True)
        for b in range(len(self)):
            idx = order[b * self.batch_size:(b + 1) * self.batch_size]
            p3, p2, cam = gather_pairs(self.bank.pose3d, self.bank.pose2d, self.bank.cam, idx, rows=n)
            yield p3, p2, ["none"] * idx.shape[0], cam
