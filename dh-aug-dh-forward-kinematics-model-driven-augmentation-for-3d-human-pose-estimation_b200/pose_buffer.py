"""Device-resident fake-pair bank.

The reference appends every iteration's ``pos_3d_cam`` / ``uv`` / ``cam_para_temp`` to host lists
(models_Fk_GAN/model_fk_gan_train.py:486-488), concatenates them at the end of the GAN epoch and re-serves them
to ``train_posenet`` through ``DataLoader(PoseDataSet(...), batch_size, shuffle=True, pin_memory=True)``
(:504-510; common/data_loader.py:9-36) -- one D2H copy per iteration and one H2D copy per batch.

``DevicePoseBuffer`` keeps the pairs in HBM (bounded ring of 384-byte records, one per pose: a shuffled batch
reads one contiguous span per pose; 180 GB per GPU holds 4.7e8 pairs) and
``DevicePoseBuffer.loader()`` yields batches with the same wire format and -- for the same torch seed -- the SAME
sample order as the reference's shuffled DataLoader under the installed torch: the iterator draws a base seed,
RandomSampler draws its own int64 seed from torch's default RNG and permutes with ``torch.randperm(n, generator=g)``
(torch/utils/data/{dataloader,sampler}.py), which ``shuffled_order`` reproduces draw for draw.  A mini-batch is one gather launch (dhfk_bank_gather).
"""
from __future__ import annotations

import torch

from . import _cabi
from .functional import _on_device, _require_cuda, _stream_ptr


def shuffled_order(n: int) -> torch.Tensor:
    """The permutation a `DataLoader(dataset_of_n, shuffle=True, num_workers=0)` walks in its next epoch.  Consumes
    torch's default CPU RNG exactly as torch >= 1.x does when such a loader is iterated: one int64 `random_()` draw
    for the iterator's base seed (_BaseDataLoaderIter.__init__), then one for RandomSampler's private generator."""
    torch.empty((), dtype=torch.int64).random_()                      # DataLoader iterator base seed (unused here)
    seed = int(torch.empty((), dtype=torch.int64).random_().item())   # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


REC_3D, REC_2D = 48, 80          # record layout (floats): pose3d [0,48) | pose2d [48,80) | cam [80, 80+cam_cols) | pad


def record_floats(cam_cols: int) -> int:
    """Floats per bank record: 80 + cam_cols rounded up so that a record is a whole number of 128-byte lines
    (9 camera columns -> 96 floats = 384 bytes)."""
    return (80 + int(cam_cols) + 31) // 32 * 32


def gather_pairs(records, idx, cam_cols=9, rows=None, want_cam=True):
    """(pose3d [nb,16,3], pose2d [nb,16,2], cam [nb,cam_cols]) = the records `idx` of the bank, in one launch.
    records: [capacity, rec_floats] float32 CUDA, contiguous; idx int64 on the device; `rows` = valid bank rows."""
    _require_cuda()
    lib = _cabi.load()
    if not (records.is_cuda and records.dtype == torch.float32 and records.is_contiguous() and records.dim() == 2):
        raise ValueError("records must be a contiguous 2-D float32 CUDA tensor")
    device = records.device
    idx = idx.to(device=device, dtype=torch.int64).contiguous()
    nb = idx.shape[0]
    rows = records.shape[0] if rows is None else int(rows)
    o3 = torch.empty((nb, 16, 3), dtype=torch.float32, device=device)
    o2 = torch.empty((nb, 16, 2), dtype=torch.float32, device=device)
    oc = torch.empty((nb, cam_cols), dtype=torch.float32, device=device) if want_cam else None
    with _on_device(device):
        rc = lib.dhfk_bank_gather(records.data_ptr(), records.shape[1], cam_cols, idx.data_ptr(), nb, rows,
                                  o3.data_ptr(), o2.data_ptr(), oc.data_ptr() if oc is not None else None,
                                  _stream_ptr(device))
    _cabi.check(rc, "dhfk_bank_gather")
    return o3, o2, oc


class DevicePoseBuffer:
    """Append-only (ring when full) bank of (pose3d_cam [16,3], pose2d [16,2], cam [cam_cols]) records in HBM.
    `pose3d` / `pose2d` / `cam` are strided views into the record array."""

    def __init__(self, capacity: int, device=None, cam_cols: int = 9):
        _require_cuda()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.capacity, self.cam_cols = int(capacity), int(cam_cols)
        self.records = torch.zeros((self.capacity, record_floats(cam_cols)), dtype=torch.float32, device=self.device)
        self.pose3d = self.records[:, :REC_3D].unflatten(1, (16, 3))
        self.pose2d = self.records[:, REC_3D:REC_2D].unflatten(1, (16, 2))
        self.cam = self.records[:, REC_2D:REC_2D + self.cam_cols]
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
            p3, p2, cam = gather_pairs(self.bank.records, idx, self.bank.cam_cols, rows=n)
            yield p3, p2, ["none"] * idx.shape[0], cam
