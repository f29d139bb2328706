"""SURVEY 8 f4 -- device-resident fake-pair bank (models_Fk_GAN/model_fk_gan_train.py:486-510;
common/data_loader.py:9-36 PoseDataSet).  Index work is bit-exact: the shuffled order equals the one the
reference's `DataLoader(..., shuffle=True)` walks for the same torch seed, and a served mini-batch equals
fancy indexing of the concatenated buffers."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, Dataset


class _PoseDataSet(Dataset):
    """Restatement of common/data_loader.py:9-36 (lists of per-iteration numpy buffers -> concatenated arrays)."""

    def __init__(self, poses_3d, poses_2d, actions, cams):
        self._p3, self._p2, self._cams = np.concatenate(poses_3d), np.concatenate(poses_2d), np.concatenate(cams)
        self._actions = sum(actions, [])

    def __getitem__(self, i):
        return torch.from_numpy(self._p3[i]).float(), torch.from_numpy(self._p2[i]).float(), self._actions[i], self._cams[i]

    def __len__(self):
        return len(self._actions)


@pytest.mark.parametrize("n,bs", [(1, 1), (7, 3), (1000, 64), (4608, 512)])
def test_shuffled_order_equals_dataloader(n, bs):
    from dhfk import pose_buffer as pb
    torch.manual_seed(1234)
    dl = DataLoader(torch.arange(n), batch_size=bs, shuffle=True)
    want = torch.cat([b for b in dl])
    after = torch.rand(3)
    torch.manual_seed(1234)
    got = pb.shuffled_order(n)
    assert torch.equal(got, want)
    assert torch.equal(torch.rand(3), after)          # same RNG position afterwards
    assert sorted(got.tolist()) == list(range(n))


def _iterations(rng, iters, batch):
    p3 = [rng.randn(batch, 16, 3).astype(np.float32) for _ in range(iters)]
    p2 = [rng.randn(batch, 16, 2).astype(np.float32) for _ in range(iters)]
    cams = [np.repeat(rng.randn(1, 9).astype(np.float32), batch, 0) for _ in range(iters)]
    return p3, p2, cams


@pytest.mark.gpu
def test_gather_is_bit_exact_and_guards_indices():
    from dhfk import pose_buffer as pb
    rng = np.random.RandomState(0)
    p3, p2, cams = _iterations(rng, 1, 5000)
    bank = pb.DevicePoseBuffer(5000, device="cuda:0")
    assert bank.records.shape == (5000, 96) and pb.record_floats(9) == 96 and pb.record_floats(16) == 96 and pb.record_floats(17) == 128
    bank.append(*(torch.tensor(a[0], device="cuda:0") for a in (p3, p2, cams)))
    b3, b2, bc = (torch.tensor(a[0], device="cuda:0") for a in (p3, p2, cams))
    assert torch.equal(bank.pose3d, b3) and torch.equal(bank.pose2d, b2) and torch.equal(bank.cam, bc)
    idx = torch.tensor(rng.randint(0, 5000, 777), device="cuda:0")
    o3, o2, oc = pb.gather_pairs(bank.records, idx)
    assert torch.equal(o3, b3[idx]) and torch.equal(o2, b2[idx]) and torch.equal(oc, bc[idx])
    o3, o2, oc = pb.gather_pairs(bank.records, torch.tensor([0, 4999, 5000, -1], device="cuda:0"))
    assert torch.equal(o3[:2], b3[[0, 4999]]) and torch.isnan(o3[2:]).all() and torch.isnan(o2[2:]).all() and torch.isnan(oc[2:]).all()
    o3, o2, oc = pb.gather_pairs(bank.records, idx, rows=100)          # only the first 100 records are valid
    bad = idx >= 100
    assert torch.isnan(o3[bad]).all() and torch.equal(o3[~bad], b3[idx[~bad]])
    e3, e2, ec = pb.gather_pairs(bank.records, torch.zeros(0, dtype=torch.int64, device="cuda:0"))
    assert e3.shape == (0, 16, 3) and e2.shape == (0, 16, 2) and ec.shape == (0, 9)
    n3, n2, nc = pb.gather_pairs(bank.records, idx, want_cam=False)
    assert nc is None and torch.equal(n3, b3[idx])
    wide = pb.DevicePoseBuffer(64, device="cuda:0", cam_cols=16)        # the reference's 16-column camera rows
    c16 = torch.randn(64, 16, device="cuda:0")
    wide.append(b3[:64], b2[:64], c16)
    w3, w2, wc = pb.gather_pairs(wide.records, torch.arange(63, -1, -1, device="cuda:0"), cam_cols=16)
    assert torch.equal(wc, c16.flip(0)) and torch.equal(w3, b3[:64].flip(0))


@pytest.mark.gpu
@pytest.mark.parametrize("iters,batch,bs", [(6, 100, 64), (3, 1024, 1024), (5, 33, 7)])
def test_loader_serves_what_the_reference_dataloader_serves(iters, batch, bs):
    """Same torch seed -> same batches (content and order) as DataLoader(PoseDataSet(...), shuffle=True)."""
    from dhfk import pose_buffer as pb
    rng = np.random.RandomState(iters)
    p3, p2, cams = _iterations(rng, iters, batch)
    bank = pb.DevicePoseBuffer(capacity=iters * batch, device="cuda:0")
    for a, b, c in zip(p3, p2, cams):
        bank.append(torch.tensor(a, device="cuda:0"), torch.tensor(b, device="cuda:0"), torch.tensor(c, device="cuda:0"))
    assert len(bank) == iters * batch
    ref = DataLoader(_PoseDataSet(p3, p2, [["none"] * (iters * batch)], cams), batch_size=bs, shuffle=True)
    for epoch in range(2):
        torch.manual_seed(77 + epoch)
        want = [(a.clone(), b.clone(), c.clone()) for a, b, _, c in ref]
        torch.manual_seed(77 + epoch)
        loader = bank.loader(bs)
        assert len(loader) == len(ref)
        got = list(loader)
        assert len(got) == len(want)
        for (g3, g2, act, gc), (w3, w2, wc) in zip(got, want):
            assert g3.is_cuda and torch.equal(g3.cpu(), w3) and torch.equal(g2.cpu(), w2) and torch.equal(gc.cpu(), wc)
            assert act == ["none"] * g3.shape[0]


@pytest.mark.gpu
def test_ring_overwrites_oldest_rows():
    from dhfk import pose_buffer as pb
    bank = pb.DevicePoseBuffer(capacity=10, device="cuda:0")
    mk = lambda lo, hi: torch.arange(lo, hi, dtype=torch.float32, device="cuda:0")
    for lo, hi in ((0, 4), (4, 8), (8, 13)):             # 13 rows into a ring of 10: rows 3..12 survive
        v = mk(lo, hi)
        bank.append(v.view(-1, 1, 1).expand(-1, 16, 3), v.view(-1, 1, 1).expand(-1, 16, 2), v.view(-1, 1).expand(-1, 9))
    assert len(bank) == 10 and bank.count == 13
    assert sorted(bank.pose3d[:, 0, 0].tolist()) == list(range(3, 13))
    assert torch.equal(bank.pose3d[:, 0, 0], bank.pose2d[:, 0, 0]) and torch.equal(bank.pose3d[:, 0, 0], bank.cam[:, 0])
    big = mk(100, 125)
    bank.append(big.view(-1, 1, 1).expand(-1, 16, 3), big.view(-1, 1, 1).expand(-1, 16, 2), big.view(-1, 1).expand(-1, 9))
    assert sorted(bank.pose3d[:, 0, 0].tolist()) == list(range(115, 125))
    bank.reset()
    assert len(bank) == 0 and len(bank.loader(4)) == 0 and list(bank.loader(4)) == []
