#!/usr/bin/env python
"""BASELINE configs[2] as a measurement: one iteration of the single-frame DH-AUG GAN loop (Gen / Dis dense 256,
batch 1024; models_Fk_GAN/model_fk_gan_train.py:287-510) on one GPU, two ways:

  native  the drop-in pieces of this repo: fused generator epilogue + FK (dhfk.Fk_generator.Fk_Generator), fused
          world->camera / projection, fused critic inputs (root-centring, flip, KCS with WGAN-GP double backward),
          device-resident fake-pair bank
  eager   the reference's own op sequence in torch eager on the same GPU (oracle/torch_port.py: per-entry DH matrix
          writes, bmm chains, column scatters, qrot, KCS row writes), host-copy fake-pair lists

The MLPs (cuBLAS) are identical and share their weights.  An iteration = 3-D critic step (+ flip pass), 2-D critic step
(+ flip pass), generator step, fake-pair append -- the reference does the generator step every 5th iteration; it is
done every iteration here so that both arms time all three.  Prints one JSON object."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dhfk  # noqa: E402
import torch_port  # noqa: E402
from dhfk import Fk_discriminator as native_dis  # noqa: E402
from dhfk import Fk_generator as native_gen  # noqa: E402
from dhfk import camera as native_cam  # noqa: E402
from dhfk import pose_buffer, synthetic, tables  # noqa: E402

LEFT, RIGHT = [4, 5, 6, 10, 11, 12], [1, 2, 3, 13, 14, 15]
ZERO_SLOTS = (4, 9, 22, 23, 28, 33)
IDX16 = [0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27]


class Res(nn.Module):                         # special_operate.py:490-510
    def __init__(self, d):
        super().__init__()
        self.fc1, self.fc2, self.relu = nn.Linear(d, d), nn.Linear(d, d), nn.ReLU(True)

    def forward(self, x):
        out = self.fc2(self.relu(self.fc1(x)))
        out += x
        return self.relu(out)


class Critic3D(nn.Module):                    # Fk_discriminator.py:149-206
    def __init__(self, d, kcs):
        super().__init__()
        self.kcs = kcs
        self.previous = nn.Sequential(nn.Linear(48, d), nn.ReLU(True))
        self.block1, self.block2, self.block3 = Res(d), Res(d), Res(d)
        self.special_KCS_previous = nn.Sequential(nn.Linear(30, d), nn.ReLU(True))
        self.special_KCS_block1, self.special_KCS_block2, self.special_KCS_block3 = Res(d), Res(d), Res(d)
        self.merge_previous = nn.Sequential(nn.Linear(2 * d, 100), nn.ReLU(True))
        self.merge_block1 = Res(100)
        self.output = nn.Linear(100, 1)

    def forward(self, inp):
        k = self.kcs(torch.clone(inp)).contiguous().view(-1, 30)
        k = self.special_KCS_block3(self.special_KCS_block2(self.special_KCS_block1(self.special_KCS_previous(k))))
        p = self.block3(self.block2(self.block1(self.previous(inp.contiguous().view(-1, 48)))))
        return self.output(self.merge_block1(self.merge_previous(torch.cat((k, p), dim=-1))))


class Critic2D(nn.Module):                    # Fk_discriminator.py:238-266
    def __init__(self, d):
        super().__init__()
        self.l1, self.l2, self.l3, self.l4 = nn.Linear(32, d), nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
        self.last, self.pred, self.relu = nn.Linear(d, d), nn.Linear(d, 1), nn.LeakyReLU()

    def forward(self, x):
        x = x.contiguous().view(-1, 32)
        d1 = self.relu(self.l1(x))
        d3 = self.relu(self.l3(self.relu(self.l2(d1))) + d1)
        return self.pred(self.relu(self.last(self.l4(d3))))


def gradient_penalty(D, real, fake, lam=10.0):          # Fk_discriminator.py:208-233
    b = real.shape[0]
    alpha = torch.rand(b, 1, device=real.device).expand(b, real[0].numel())
    inter = (alpha * real.reshape(b, -1) + (1 - alpha) * fake.reshape(b, -1)).detach().requires_grad_(True)
    out = D(inter)
    (g,) = torch.autograd.grad(out, inter, grad_outputs=torch.ones_like(out), create_graph=True, retain_graph=True)
    return ((g.norm(2, dim=1) - 1) ** 2).mean() * lam


ALLREDUCE = [None]      # set under torchrun: dhfk.parallel.allreduce_grads_flat (one flat NCCL all-reduce per model step)


def critic_step(D, opt, real, fake):                     # model_fk_gan_train.py:177-232
    D.zero_grad(set_to_none=True)
    (-D(real).mean()).backward()
    D(fake).mean().backward()
    gradient_penalty(D, real.data, fake.data).backward()
    if ALLREDUCE[0] is not None:
        ALLREDUCE[0](list(D.parameters()))
    opt.step()


def flip_eager(x):                                       # model_fk_gan_train.py:324-327
    y = x.detach().clone()
    y[:, :, 0] *= -1
    y[:, LEFT + RIGHT, :] = y[:, RIGHT + LEFT, :]
    return y


class EagerGenerator(nn.Module):
    """Fk_generator.py:114-259 with the reference's op sequence: MLP, two tanh, 37 column copies + 37 scaled column
    writes, 15 length products, the 34-matrix FK (torch_port), the 32->16 gather."""

    def __init__(self, mlp_owner, bone):
        super().__init__()
        self.owner, self.bone = mlp_owner, bone
        rng = np.concatenate([tables.GAN_ANGLE_RANGE, tables.GAN_GLOBAL_ROT_RANGE]).astype(np.float32)
        self.lo, self.hi = rng[:, 0], rng[:, 1]

    def forward(self, noise):
        out = self.owner._mlp(noise)
        out = torch.cat((torch.tanh(out[:, :-3]), torch.tanh(out[:, -3:]) * 10.0), dim=1)
        n = out.shape[0]
        g = torch.zeros((n, 37), dtype=torch.float32, device=out.device)
        k = 0
        for i in range(37):
            if i not in ZERO_SLOTS:
                g[:, i] = out[:, k]
                k += 1
        for i in range(37):
            g[:, i] = g[:, i] * float((self.hi[i] - self.lo[i]) / 2) + float((self.hi[i] + self.lo[i]) / 2)
        scaler = torch.randint(-200, 200, size=(n, 8), device=out.device) / 1000.0
        bone = native_gen.scaled_bone_lengths(self.bone, scaler)
        w32 = torch_port.RefFKPort(n, out.device).fk32(g[:, :33], g[:, 34:37], bone, out[:, -3:])
        return w32[:, IDX16].reshape(n, 48)


def make_arm(kind, G_native, bone, dense, dev, state):
    if kind == "native":
        G = G_native
        kcs = lambda x: native_dis.special_KCS_Input_transform(x)
        w2c, proj = native_cam.GAN_torch_world_to_camera, native_cam.project_to_2d
        centre = lambda x: dhfk.critic_input(x, centre=True, kcs_cols=0)
        flip3 = lambda x: dhfk.critic_input(x.detach(), flip=True, kcs_cols=0)
        flip2 = lambda x: dhfk.flip_pose(x.detach())
    else:
        G = EagerGenerator(G_native, bone)
        kcs = torch_port.special_kcs
        w2c, proj = (lambda x, R, t: torch_port.world_to_camera(x, R, t)), torch_port.project_to_2d
        centre = lambda x: x[:, :, :] - x[:, :1, :]
        flip3 = flip2 = flip_eager
    D3, D2 = Critic3D(dense, kcs).to(dev), Critic2D(dense).to(dev)
    D3.load_state_dict(state["d3"]); D2.load_state_dict(state["d2"])
    return dict(G=G, D3=D3, D2=D2, w2c=w2c, proj=proj, centre=centre, flip3=flip3, flip2=flip2,
                o3=torch.optim.Adam(D3.parameters(), 1e-4, capturable=True),
                o2=torch.optim.Adam(D2.parameters(), 1e-4, capturable=True))


def iteration(arm, kind, oG, real3d, real2d, cam_q, cam_t, cam_rows, bank, host_lists, batch, dev, append=True):
    G, D3, D2 = arm["G"], arm["D3"], arm["D2"]
    with torch.no_grad():
        fake = G(torch.randn(batch, 128, device=dev)).view(-1, 16, 3)
    fake_c = arm["centre"](fake)
    critic_step(D3, arm["o3"], real3d, fake_c)
    critic_step(D3, arm["o3"], arm["flip3"](real3d), arm["flip3"](fake_c))
    cam = arm["w2c"](fake, cam_q, cam_t)
    uv = arm["proj"](cam, cam_rows)
    critic_step(D2, arm["o2"], real2d, uv)
    critic_step(D2, arm["o2"], arm["flip2"](real2d), arm["flip2"](uv))
    # generator step (model_fk_gan_train.py:415-482)
    for p in list(D3.parameters()) + list(D2.parameters()):
        p.requires_grad_(False)
    oG.zero_grad(set_to_none=True)
    fake = G(torch.randn(batch, 128, device=dev)).view(-1, 16, 3)
    uv_g = arm["proj"](arm["w2c"](fake, cam_q, cam_t), cam_rows)
    fc = arm["centre"](fake)
    loss = (D3(fc).mean() + D3(arm["flip3"](fc)).mean()) / 2 + 0.2 * (D2(uv_g).mean() + D2(arm["flip2"](uv_g)).mean()) / 2
    (-loss).backward()
    if ALLREDUCE[0] is not None:
        ALLREDUCE[0](list(G.parameters()))
    oG.step()
    for p in list(D3.parameters()) + list(D2.parameters()):
        p.requires_grad_(True)
    # fake-pair buffer (model_fk_gan_train.py:486-488)
    if not append:
        return cam, uv
    if kind == "native":
        bank.append(cam, uv, cam_rows)
    else:
        host_lists.append((cam.detach().cpu().numpy(), uv.detach().cpu().numpy(), cam_rows.detach().cpu().numpy()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--dense", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--eager-iters", type=int, default=5)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:       # data-parallel GAN iteration (BASELINE configs[4]): every rank its own batch, grads all-reduced
        import torch.distributed as dist
        from dhfk import parallel
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        ALLREDUCE[0] = parallel.allreduce_grads_flat
    torch.manual_seed(0)
    args = argparse.Namespace(batch_size=a.batch, random_seed=0, single_or_multi_train_mode="single", architecture="3,3,3",
                              GAN_OUTPUT_DIM=35, Gen_DenseDim=a.dense, GAN_whether_use_preAngle=True, whether_use_RT=True,
                              bone_len_scaler="different", record_all_picture=False, checkpoint="/tmp")
    fk = dhfk.Forward_Kinematics_DH_Model(args, ["S1"], None)
    G = native_gen.Fk_Generator(fk, args, dev).to(dev)
    with torch.no_grad():                          # keep the fakes in front of the camera
        G.deconv_out.weight.mul_(0.05)
        G.deconv_out.bias[-3:] = torch.tensor([0.0, 0.0, 0.1])
    state = {"d3": Critic3D(a.dense, None).state_dict(), "d2": Critic2D(a.dense).state_dict()}
    torch.manual_seed(1 + rank)                    # same weights everywhere, different noise / data per rank
    inp = synthetic.gan_like(a.batch, seed=3 + rank)
    bone = torch.tensor(inp["bone"], device=dev)
    G.boneLength = bone
    blk = tables.camera_block("S1", 0)
    cam_q = torch.tensor(blk[0:4], device=dev).view(1, 4)
    cam_t = torch.tensor(blk[4:7], device=dev).view(1, 3)
    cam_rows = torch.tensor(blk[7:16], device=dev).view(1, 9).repeat(a.batch, 1)
    w, _, uv = dhfk.fk_project(*(torch.tensor(inp[k], device=dev) for k in ("ang", "grot", "bone", "root")), blk, return_cam=False)
    real3d = (w - w[:, :1]).detach()
    real2d = uv.detach()
    out = {"config": "BASELINE configs[2]: single-frame DH-AUG GAN iteration, Gen/Dis dense %d, batch %d" % (a.dense, a.batch),
           "iteration": "D3D step + flip pass, D2D step + flip pass (WGAN-GP each), generator step, fake-pair append"}
    if world > 1:
        out["data_parallel"] = "%d ranks x batch %d, one flat NCCL all-reduce per optimiser step (5 per iteration)" % (world, a.batch)
    for kind, iters in ((("native", a.iters),) if world > 1 else (("native", a.iters), ("eager", a.eager_iters))):
        arm = make_arm(kind, G, bone, a.dense, dev, state)
        oG = torch.optim.Adam(G.parameters(), 1e-4)
        bank = pose_buffer.DevicePoseBuffer(a.batch * (iters + 2), device=dev)
        host_lists = []
        for _ in range(2):
            iteration(arm, kind, oG, real3d, real2d, cam_q, cam_t, cam_rows, bank, host_lists, a.batch, dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            iteration(arm, kind, oG, real3d, real2d, cam_q, cam_t, cam_rows, bank, host_lists, a.batch, dev)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / iters * 1e3
        out[kind] = {"ms_per_iteration": ms, "poses_per_s": 2 * a.batch * world / (ms * 1e-3), "iterations": iters}
    if world > 1:
        import torch.distributed as dist
        # replicas must still agree after the all-reduced steps
        flat = torch.cat([p.detach().reshape(-1) for m in (G, arm["D3"], arm["D2"]) for p in m.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        out["replicas_in_sync"] = bool(torch.equal(flat, ref))
        dist.barrier()
        if rank == 0:
            print(json.dumps(out, indent=1))
        dist.destroy_process_group()
        return
    # native pieces with the whole iteration captured in ONE CUDA graph (the C-ABI launches are plain stream work, so
    # torch.cuda.graph captures them with the cuBLAS kernels; only the bank append stays outside: it moves a host-side
    # ring head).  Scalers are drawn on the device (Fk_Generator.scaler_source) -- no host RNG inside the capture.
    try:
        G.scaler_source = lambda rows, frames: torch.randint(-200, 200, (rows * frames, 8), device=dev) / 1000.0
        arm = make_arm("native", G, bone, a.dense, dev, state)
        oG = torch.optim.Adam(G.parameters(), 1e-4, capturable=True)
        bank = pose_buffer.DevicePoseBuffer(a.batch * (a.iters + 2), device=dev)
        core = lambda: iteration(arm, "native", oG, real3d, real2d, cam_q, cam_t, cam_rows, bank, None, a.batch, dev, append=False)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                core()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            cam_s, uv_s = core()
        for _ in range(2):
            graph.replay(); bank.append(cam_s, uv_s, cam_rows)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.iters):
            graph.replay(); bank.append(cam_s, uv_s, cam_rows)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / a.iters * 1e3
        out["native_cuda_graph"] = {"ms_per_iteration": ms, "poses_per_s": 2 * a.batch / (ms * 1e-3), "iterations": a.iters}
    except Exception as e:                      # context only: never let it break the two measured arms
        out["native_cuda_graph"] = {"error": repr(e)[:300]}
    finally:
        G.scaler_source = None
    out["speedup"] = out["eager"]["ms_per_iteration"] / out["native"]["ms_per_iteration"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
