"""SURVEY 8 f3 -- bone-length retarget + per-row projection (random_bl_aug + project_to_2d of the per-epoch
loader refresh, function_aug/dataloader_update.py:18-41,69; video_mode_operate.py:879-928).

CPU part: the C oracle reproduces the goldens frozen from the unmodified reference.
GPU part: the fused kernel (through the public API -> ctypes -> C ABI) against goldens and oracle.
Tolerance |x - ref| <= 1e-5 * max(|ref|, 1) (north_star)."""
import numpy as np
import pytest
import torch

from conftest import assert_parity


def test_c_oracle_retarget_matches_reference(golden, c_oracle):
    g = golden("retarget")
    o = c_oracle.retarget(g["pose"], g["templates"], g["tmpl_idx"], g["cam_rows16"])
    assert_parity(o["pose"], g["out_pose"], "out_pose")
    assert_parity(o["uv"], g["out_uv"], "out_uv")
    v = c_oracle.retarget(g["v_pose"], g["templates"][g["v_idx"]], None, g["v_cam_row"][None])
    assert_parity(v["pose"], g["v_out_pose"], "v_out_pose")
    assert_parity(v["uv"], g["v_out_uv"], "v_out_uv")


def test_retarget_properties_on_oracle(golden, c_oracle):
    """Domain properties: bone lengths equal the template row exactly, unit directions and the root are kept,
    and the operation is idempotent for a fixed template choice."""
    g = golden("retarget")
    o = c_oracle.retarget(g["pose"], g["templates"], g["tmpl_idx"])["pose"]
    I = [0, 1, 2, 0, 4, 5, 0, 7, 8, 8, 10, 11, 8, 13, 14]
    J = list(range(1, 16))
    bl = np.linalg.norm(o[:, I] - o[:, J], axis=-1)
    assert np.abs(bl - g["templates"][g["tmpl_idx"]]).max() < 1e-7
    assert np.abs(o[:, 0] - g["pose"][:, 0]).max() == 0
    d0 = g["pose"][:, I] - g["pose"][:, J]
    d1 = o[:, I] - o[:, J]
    cos = (d0 * d1).sum(-1) / np.linalg.norm(d0, axis=-1) / np.linalg.norm(d1, axis=-1)
    assert np.abs(cos - 1).max() < 1e-6
    again = c_oracle.retarget(o.astype(np.float32), g["templates"], g["tmpl_idx"])["pose"]
    assert np.abs(again - o).max() < 1e-6


def test_random_stream_matches_reference(golden, monkeypatch):
    """random_bl_aug consumes np.random exactly like the reference: same template rows, same RNG position."""
    import dhfk.dataloader_update as du
    g = golden("retarget")
    seen = {}
    monkeypatch.setattr(du, "retarget_project", lambda x, tm, idx=None, *a, **k: seen.update(idx=idx, tm=tm) or x)
    np.random.seed(17)
    du.random_bl_aug(torch.zeros(g["pose"].shape))
    assert np.array_equal(seen["idx"], g["tmpl_idx"])
    assert np.random.randint(0, 1 << 30) == int(g["rng_after"][0])
    assert np.array_equal(np.asarray(seen["tm"]), g["templates"])
    np.random.seed(4)
    du.video_mode_random_bl_aug(torch.zeros(27, 16, 3))
    assert np.array_equal(np.asarray(seen["tm"]), g["templates"][g["v_idx"]]) and seen["idx"] is None


# ---------------------------------------------------------------------------------------- GPU
def T(x, dtype=torch.float32):
    return torch.tensor(np.asarray(x), dtype=dtype, device="cuda:0")


@pytest.mark.gpu
def test_kernel_matches_reference_golden(golden):
    import dhfk
    g = golden("retarget")
    pose, uv = dhfk.functional.retarget_project(T(g["pose"]), g["templates"], g["tmpl_idx"], T(g["cam_rows16"]))
    assert_parity(pose.cpu().numpy(), g["out_pose"], "out_pose")
    assert_parity(uv.cpu().numpy(), g["out_uv"], "out_uv")
    only = dhfk.functional.retarget_project(T(g["pose"]), g["templates"], g["tmpl_idx"])
    assert torch.equal(only, pose)
    # sequence variant: one template row, one shared intrinsics row (stride 0)
    vp, vuv = dhfk.functional.retarget_project(T(g["v_pose"]), g["templates"][g["v_idx"]], None, T(g["v_cam_row"]))
    assert_parity(vp.cpu().numpy(), g["v_out_pose"], "v_out_pose")
    assert_parity(vuv.cpu().numpy(), g["v_out_uv"], "v_out_uv")


@pytest.mark.gpu
def test_reference_shaped_random_bl_aug(golden):
    import dhfk.dataloader_update as du
    g = golden("retarget")
    np.random.seed(17)
    out = du.random_bl_aug(T(g["pose"]))
    assert_parity(out.cpu().numpy(), g["out_pose"], "random_bl_aug")
    assert np.random.randint(0, 1 << 30) == int(g["rng_after"][0])
    np.random.seed(4)
    vout = du.video_mode_random_bl_aug(T(g["v_pose"]))
    assert_parity(vout.cpu().numpy(), g["v_out_pose"], "video_mode_random_bl_aug")
    # CPU input is accepted like every other entry point (moved to the device; no CPU compute path)
    np.random.seed(17)
    out2 = du.random_bl_aug(torch.tensor(g["pose"]))
    assert out2.is_cuda and torch.equal(out2, out)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1000, 262144 + 5])
def test_kernel_matches_oracle_ragged_and_large(c_oracle, n):
    import dhfk
    from dhfk import synthetic, tables
    rng = np.random.RandomState(100 + n % 97)
    d = synthetic.gan_like(max(n, 1), seed=7)
    o = c_oracle.forward(d["ang"], d["grot"], d["bone"], d["root"], tables.camera_block("S5", 1))
    pose = o["cam"].astype(np.float32)[:n]
    idx = rng.randint(0, 5, n).astype(np.int32)
    rows = np.stack([tables.camera_block(tables.TRAIN_SUBJECTS[rng.randint(5)], rng.randint(4))[7:16]
                     for _ in range(8)])[rng.randint(0, 8, n)].reshape(n, 9)
    got_p, got_uv = dhfk.functional.retarget_project(T(pose), tables.BONE_TEMPLATES_GANUTILS_ORDER, idx, T(rows))
    assert got_p.shape == (n, 16, 3) and got_uv.shape == (n, 16, 2)
    if n == 0:
        return
    ref = c_oracle.retarget(pose, tables.BONE_TEMPLATES_GANUTILS_ORDER, idx, rows)
    assert_parity(got_p.cpu().numpy(), ref["pose"], "pose")
    assert_parity(got_uv.cpu().numpy(), ref["uv"], "uv")
    # size-independent properties at full size: lengths are the template's, root untouched, in-place aliasing
    I = [0, 1, 2, 0, 4, 5, 0, 7, 8, 8, 10, 11, 8, 13, 14]
    bl = (got_p[:, I] - got_p[:, 1:]).norm(dim=-1).cpu().numpy()
    assert np.abs(bl - tables.BONE_TEMPLATES_GANUTILS_ORDER[idx]).max() < 2e-6
    assert torch.equal(got_p[:, 0], T(pose)[:, 0])
    buf = T(pose)
    dhfk.functional.retarget_project(buf, tables.BONE_TEMPLATES_GANUTILS_ORDER, idx, T(rows), out_pose=buf)
    assert torch.equal(buf, got_p)


@pytest.mark.gpu
def test_degenerate_bone_gives_nan_like_reference():
    """A zero-length bone is 0/0 in the reference (gan_utils.py:133) -> NaN below it in the tree; same here."""
    import dhfk
    from dhfk import tables
    pose = np.random.RandomState(0).randn(4, 16, 3).astype(np.float32)
    pose[1, 2] = pose[1, 1]          # bone 1 (joint 1 -> 2) of pose 1 collapses
    out = dhfk.functional.retarget_project(T(pose), tables.BONE_TEMPLATES_GANUTILS_ORDER, np.zeros(4, np.int32))
    out = out.cpu().numpy()
    assert np.isnan(out[1, 2]).all() and np.isnan(out[1, 3]).all()
    assert np.isfinite(out[1, [0, 1, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]]).all()
    assert np.isfinite(out[[0, 2, 3]]).all()


# ---- the two loader-refresh entry points with stand-ins for the reference's own Dataset / generator classes ---------
class _PoseDataSet:                      # common/data_loader.py:9-36
    def __init__(self, poses_3d, poses_2d, actions, cams):
        self._p3, self._p2, self._cams = np.concatenate(poses_3d), np.concatenate(poses_2d), np.concatenate(cams)
        self._actions = sum((list(a) for a in actions), [])

    def __getitem__(self, i):
        return torch.from_numpy(self._p3[i]).float(), torch.from_numpy(self._p2[i]).float(), self._actions[i], self._cams[i]

    def __len__(self):
        return len(self._actions)


class _PoseTarget:                       # common/data_loader.py:62-74
    def __init__(self, poses):
        self._poses = np.concatenate(poses)

    def __getitem__(self, i):
        return torch.from_numpy(self._poses[i]).float()

    def __len__(self):
        return len(self._poses)


@pytest.mark.gpu
def test_dataloader_update_contract(c_oracle, monkeypatch):
    import argparse
    import sys
    import types
    from torch.utils.data import DataLoader
    import dhfk.dataloader_update as du
    from dhfk import tables
    mod = types.ModuleType("common.data_loader")
    mod.PoseDataSet, mod.PoseTarget = _PoseDataSet, _PoseTarget
    monkeypatch.setitem(sys.modules, "common", types.ModuleType("common"))
    monkeypatch.setitem(sys.modules, "common.data_loader", mod)
    rng = np.random.RandomState(3)
    n = 300
    p3 = rng.randn(n, 16, 3).astype(np.float32) * 0.3 + np.array([0, 0, 5], np.float32)
    cams = np.stack([tables.camera_block(tables.TRAIN_SUBJECTS[rng.randint(5)], rng.randint(4))[7:16] for _ in range(n)])
    acts = ["act%d" % (i % 7) for i in range(n)]
    data_dict = {"train_gt2d3d_loader": DataLoader(_PoseDataSet([p3], [np.zeros((n, 16, 2), np.float32)], [acts], [cams]),
                                                   batch_size=128, shuffle=False)}
    args = argparse.Namespace(batch_size=64, num_workers=0)
    np.random.seed(5)
    du.dataloader_update(args, data_dict, torch.device("cuda:0"))
    np.random.seed(5)
    idx = np.concatenate([np.random.choice(5, k) for k in (128, 128, 44)])      # one choice() per batch, like the reference
    ref = c_oracle.retarget(p3, tables.BONE_TEMPLATES_GANUTILS_ORDER, idx, cams)
    ds = data_dict["train_gt2d3d_loader"].dataset
    assert_parity(ds._p3, ref["pose"], "refreshed poses")
    assert_parity(ds._p2, ref["uv"], "refreshed 2-D")
    assert np.array_equal(ds._cams, cams) and ds._actions == acts
    assert np.array_equal(data_dict["target_3d_loader"].dataset._poses, ds._p3)
    assert np.array_equal(data_dict["target_2d_loader"].dataset._poses, ds._p2)
    assert data_dict["train_gt2d3d_loader"].batch_size == 64


@pytest.mark.gpu
def test_video_mode_dataloader_update_contract(c_oracle, monkeypatch):
    import argparse
    import sys
    import types
    import dhfk.dataloader_update as du
    from dhfk import tables
    made = {}

    class Gen:                           # stands in for GAN_video_ChunkedGenerator (video_mode_operate.py:35)
        def __init__(self, batch_size, cameras, poses_3d, poses_2d, **kw):
            made.update(batch_size=batch_size, cameras=cameras, poses_3d=poses_3d, poses_2d=poses_2d, kw=kw)

    mod = types.ModuleType("models_Fk_GAN.video_mode_operate")
    mod.GAN_video_ChunkedGenerator = Gen
    mod.video_receptive_field = lambda fw: int(np.prod(fw))
    monkeypatch.setitem(sys.modules, "models_Fk_GAN", types.ModuleType("models_Fk_GAN"))
    monkeypatch.setitem(sys.modules, "models_Fk_GAN.video_mode_operate", mod)
    rng = np.random.RandomState(8)
    lens = [57, 1, 200, 33]
    seqs = [rng.randn(k, 16, 3).astype(np.float32) * 0.3 + np.array([0, 0, 4.5], np.float32) for k in lens]
    cams = [np.concatenate([tables.camera_block("S%d" % s, c)[7:16], rng.randn(7).astype(np.float32)])
            for s, c in ((1, 0), (5, 1), (6, 2), (8, 3))]                        # 16-column camera rows, first 9 used
    data_dict = {"poses_train": seqs, "poses_train_2d": [None] * 4, "actions_train": ["a", "b", "c", "d"], "cams_train": cams}
    args = argparse.Namespace(batch_size=32, architecture="3,3,3")
    np.random.seed(21)
    du.video_mode_dataloader_update(args, data_dict, torch.device("cuda:0"))
    np.random.seed(21)
    picks = [np.random.choice(5, 1)[0] for _ in lens]                            # one row per sequence
    assert made["batch_size"] == 32 and made["kw"]["pad"] == 13 and made["kw"]["shuffle"] is True
    assert [p.shape[0] for p in made["poses_3d"]] == lens
    for seq, cam, pick, got3, got2, gotc in zip(seqs, cams, picks, made["poses_3d"], made["poses_2d"], made["cameras"]):
        ref = c_oracle.retarget(seq, tables.BONE_TEMPLATES_GANUTILS_ORDER[[pick]], None, cam[None, :9])
        assert_parity(got3, ref["pose"], "sequence poses")
        assert_parity(got2, ref["uv"], "sequence 2-D")
        assert np.array_equal(gotc, cam)
