"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference's own GAN training loops on synthetic data.

  GAN_solutions_FK_generator              models_Fk_GAN/model_fk_gan_train.py:236-512   (BASELINE configs[2])
  video_mode_GAN_solutions_FK_generator   models_Fk_GAN/video_GAN_fun.py:79-602         (BASELINE configs[3])

The reference is imported from /root/reference (build container) or from the archive oracle/stage_ref.py made
(GPU box).  Two ways to run a loop:

  device="cpu"   nothing patched except `torch.device("cuda")` inside the two loop modules, which is redirected to the
                 CPU (the loops hard-code it) -- this is the golden: the reference's own FK, camera, critics, on CPU;
  device="cuda"  on a GPU, after `dhfk.dropin.install(...)`: the same unmodified loop functions, with the native kernels
                 underneath.  Never used by the product; tests and bench.py's reference legs only.

What is recorded: every scalar the loop hands to its TensorBoard writer (D_real / D_fake / Wasserstein distance of
every critic step, per iteration), the generator's gradients at every generator step (read just before
`optimizer_G.step()`), and the fake-pair buffer the loop leaves behind.  RNG: the loops draw noise, GP alphas and bone
scalers from torch's CPU generator and cameras from np.random, so a CPU run and a GPU run see the same draws.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn

import ref_harness as rh

_MODS = {}


def load(force_cpu: bool, prefer_staged: bool = False):
    """Import the reference loop modules (once).  Visualisation helpers are replaced by no-ops: they draw with
    matplotlib (stubbed here) and write png/mp4 files; nothing on the path reads what they produce."""
    if _MODS:
        return _MODS
    ref = rh.import_reference(force_cpu=force_cpu, prefer_staged=prefer_staged)
    for name in ("tensorboardX",):
        if name not in sys.modules:
            rh._stub_module(name)
    from models_Fk_GAN import model_fk_gan_train as train          # noqa: E402
    from models_Fk_GAN import video_GAN_fun as video               # noqa: E402
    from models_Fk_GAN import special_operate                      # noqa: E402
    from models_Fk_GAN import Fk_discriminator as dis              # noqa: E402
    from function_aug import config as cfg                         # noqa: E402
    from utils import utils as uutils                              # noqa: E402
    nop = lambda *a, **k: None
    special_operate.my_draw_DOF_angle_distribute = nop             # heat-map dump every 500 generator calls
    video.my_visual_GAN_video = nop                                # mp4 of the last batch at the end of an epoch
    _MODS.update(ref=ref, train=train, video=video, special=special_operate, dis=dis, cfg=cfg, utils=uutils)
    return _MODS


class _TorchOnCPU(types.ModuleType):
    """`torch` as the loop modules see it in the golden run: identical, except that torch.device("cuda") is the CPU."""

    def __init__(self):
        super().__init__("torch")

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(spec, *a):
        if isinstance(spec, str) and spec.startswith("cuda"):
            return torch.device("cpu")
        return torch.device(spec, *a)


def redirect_cuda_to_cpu(mods):
    proxy = _TorchOnCPU()
    mods["train"].torch = proxy
    mods["video"].torch = proxy


def parse_args(mods, **over):
    """The reference's own argument parser (function_aug/config.py:5-196) with its defaults, then overrides."""
    argv, sys.argv = sys.argv, ["run_Fk_GAN.py"]
    try:
        args = mods["cfg"].get_parse_args()
    finally:
        sys.argv = argv
    for k, v in over.items():
        if not hasattr(args, k):
            raise KeyError(k)
        setattr(args, k, v)
    return args


class Recorder:
    """Stands in for tensorboardX.SummaryWriter: keeps what the loop reports."""

    def __init__(self):
        self.scalars = []

    def add_scalar(self, name, value, step):
        self.scalars.append((name, int(step), float(value)))

    def __getattr__(self, name):
        return lambda *a, **k: None


class _Skeleton:
    def num_joints(self):
        return 16


class _Dataset:
    def skeleton(self):
        return _Skeleton()


def camera_param_row(blk):
    """[f2 c2 k3 p2 | q4 | t3]: the cam_param row the loaders hand to the loops (R at 9:13, T at 13:16,
    model_fk_gan_train.py:301-302)."""
    blk = np.asarray(blk, np.float32)
    return np.concatenate([blk[7:16], blk[0:4], blk[4:7]]).astype(np.float32)


def real_batches(n_batches, rows, seed, frames=1):
    """`n_batches` batches of camera-space poses, their projections and cam_param rows (synthetic poses through the
    float64 oracle; clips are `frames` consecutive blends towards the clip's first pose)."""
    import c_oracle
    from dhfk import synthetic, tables
    blk = tables.camera_block("S1", 0)
    out = []
    for b in range(n_batches):
        inp = synthetic.gan_like(rows * frames, seed=seed + 17 * b)
        if frames > 1:      # temporally coherent clips: interpolate angles / root inside each clip
            w = np.linspace(0.0, 0.5, frames, dtype=np.float32).reshape(1, frames, 1)
            for k in ("ang", "grot", "root", "bone"):
                a = inp[k].reshape(rows, frames, -1)
                inp[k] = np.ascontiguousarray(((1 - w) * a + w * a[:, :1]).reshape(rows * frames, -1))
        o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk)
        cam3d = o["cam"].astype(np.float32)
        uv = o["uv"].astype(np.float32)
        cp = np.tile(camera_param_row(blk), (rows, 1))
        out.append((cam3d, uv, cp))
    return out


class _VideoLoader:
    """The two members of GAN_video_ChunkedGenerator the video loop touches (video_GAN_fun.py:164,166)."""

    def __init__(self, batches, rows, frames):
        self.batches, self.rows, self.frames = batches, rows, frames
        self.num_batches = len(batches)

    def next_epoch(self):
        for cam3d, uv, cp in self.batches:
            yield cp.copy(), cam3d.reshape(self.rows, self.frames, 16, 3).copy(), uv.reshape(self.rows, self.frames, 16, 2).copy()


def _grad_recorder(optimizer, params, sink):
    step = optimizer.step

    def recording_step(*a, **k):
        sink.append(np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().numpy().reshape(-1)
                                    for p in params]))
        return step(*a, **k)

    optimizer.step = recording_step


def run_loop(mode="single", *, device="cpu", iters=6, batch=32, dense=16, seed=11, install=None, prefer_staged=False,
             architecture="3,3", timers=None, fk_class=None):
    """Run the reference's loop for `iters` iterations.  mode: "single" | "video".  install: kwargs for
    dhfk.dropin.install (GPU runs) or None (the unpatched reference).  Returns dict(scalars, g_grads, params, buffer)."""
    on_cpu = device == "cpu"
    mods = load(force_cpu=on_cpu, prefer_staged=prefer_staged)
    if on_cpu:
        redirect_cuda_to_cpu(mods)
    elif install is not None:
        import dhfk
        dhfk.dropin.install(**install)
    train, video = mods["train"], mods["video"]
    multi = mode == "video"
    ckpt = tempfile.mkdtemp(prefix="dhfk_ref_loop_")
    os.makedirs(os.path.join(ckpt, "tmp"), exist_ok=True)
    args = parse_args(mods, batch_size=batch, Gen_DenseDim=dense, Dis_DenseDim_3D=dense, Dis_DenseDim_2D=dense,
                      video_Dis_DenseDim_3D=dense, video_Dis_DenseDim_2D=dense, checkpoint=ckpt, num_workers=0,
                      single_or_multi_train_mode="multi" if multi else "single", architecture=architecture,
                      record_all_picture=False, random_seed=seed)
    frames = 1
    if multi:
        for f in architecture.split(","):
            frames *= int(f)
    torch.manual_seed(seed)
    np.random.seed(seed)
    fkmod = sys.modules["models_Fk_GAN.forward_kinematics_DH_model"]
    FK = (fk_class or fkmod.Forward_Kinematics_DH_Model)(args, ["S1"], None)
    if multi:
        pose_fk = train.video_mode_my_get_poseFk_model(args, _Dataset(), FK, frames)
    else:
        pose_fk = train.my_get_poseFk_model(args, _Dataset(), FK)
    g_grads = []
    _grad_recorder(pose_fk["optimizer_G"], list(pose_fk["model_G"].parameters()), g_grads)
    batches = real_batches(iters, batch, seed + 1, frames)
    summary = mods["utils"].Summary(ckpt)
    summary.epoch = int(args.single_dis_warmup_epoch) if multi else 0      # motion critics are trained from this epoch on
    writer = Recorder()
    model_pos = nn.Linear(1, 1)
    data_dict = {}
    if multi:
        data_dict["target_GAN_loader"] = _VideoLoader(batches, batch, frames)
        fn = video.video_mode_GAN_solutions_FK_generator
    else:
        T = torch.from_numpy
        data_dict["train_gt2d3d_loader"] = [(T(c), None, None, T(cp)) for c, _, cp in batches]
        data_dict["target_2d_loader"] = [T(u) for _, u, _ in batches]
        data_dict["target_3d_loader"] = [T(c) for c, _, _ in batches]
        fn = train.GAN_solutions_FK_generator
    if timers is not None:
        timers["start"]()
    fn(args, pose_fk, data_dict, model_pos, summary, writer, ["S1"])
    if timers is not None:
        timers["stop"]()
    loader = data_dict["train_fake2d3d_loader"]
    ds = loader.dataset
    params = {name: np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in pose_fk[name].parameters()])
              for name in pose_fk if name.startswith("model_")}
    return dict(scalars=writer.scalars, g_grads=g_grads, params=params, buffer_len=len(ds), frames=frames,
                buffer_3d=np.asarray(ds._poses_3d), buffer_2d=np.asarray(ds._poses_2d))


def scalars_matrix(scalars):
    """(names, values): the recorded scalars in call order as one float64 vector, with their tags."""
    return [s[0] for s in scalars], np.array([s[2] for s in scalars], np.float64)


def time_loops(device, install, modes=("single", "video"), iters=10, batch=None, dense=256, prefer_staged=True):
    """Wall-clock ms per iteration of the reference's own loops at BASELINE configs[2] / configs[3] sizes
    (batch 1024 single-frame; 512 clips x 9 frames, architecture 3,3), after a short warm-up call of the same loop."""
    import time
    out = {}
    for mode in modes:
        b = batch or (1024 if mode == "single" else 512)
        try:
            run_loop(mode, device=device, iters=5, batch=b, dense=dense, seed=5, install=install, prefer_staged=prefer_staged)
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = run_loop(mode, device=device, iters=iters, batch=b, dense=dense, seed=6, install=install,
                         prefer_staged=prefer_staged)
            if device != "cpu":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            poses = b * r["frames"]
            out[mode] = {"ms_per_iteration": dt / iters * 1e3, "iterations": iters, "batch": b, "frames": r["frames"],
                         "dense": dense, "poses_per_iteration": poses,
                         "generator_steps": len(r["g_grads"]), "critic_steps": len(r["scalars"]) // 3}
        except Exception as e:            # e.g. the unpatched reference's CUDA branch mixing devices
            out[mode] = {"error": repr(e)[:400]}
    return out


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--install", default="all", choices=["none", "plain", "all"])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dense", type=int, default=256)
    ap.add_argument("--modes", default="single,video")
    a = ap.parse_args()
    inst = {"none": None, "plain": {}, "all": dict(generators=True, critics=True, loader_refresh=True)}[a.install]
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    fd = os.dup(1)
    os.dup2(2, 1)                       # the loops print progress bars: keep stdout for the JSON line
    res = time_loops(a.device, inst, tuple(a.modes.split(",")), a.iters, None, a.dense)
    os.write(fd, (json.dumps(res) + "\n").encode())
