#!/usr/bin/env python
"""The path an UNMODIFIED caller takes through the drop-in:

    w32  = Forward_Kinematics_DH_Model.change_3d_joint_angle(22 kwargs: slices of the generator's [N,37] tensor, 15 lengths)
    w16  = w32[:, H36M_32_To_16_Table]                       # Fk_generator.py:259
    cam  = GAN_torch_world_to_camera(w16, R, t)              # model_fk_gan_train.py:374
    uv   = project_to_2d(cam, cam_rows[N,9])                 # model_fk_gan_train.py:376
    backward through all of it (upstream gradients on w16 and uv)

timed with CUDA events at 1 M poses (roofline of the path's own kernels) and by wall clock at the reference's real
batch sizes (1 024 / 4 608: host-bound).  `python tools/dropin_path_bench.py [--profile]` prints one JSON object;
`measure(dev, ...)` is what bench.py imports."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BONES = ("left_small_leg_len", "right_small_leg_len", "left_big_leg_len", "right_big_leg_len", "left_hip_len",
         "right_hip_len", "waist_len", "thorax_len", "left_shoulder_len", "right_shoulder_len", "left_big_arm_len",
         "right_big_arm_len", "left_small_arm_len", "right_small_arm_len", "neck_len")
# bytes per pose each kernel on the path has to move (fp32): FK fwd [N,37]+bone+root -> w16; w2c; project (per-row
# intrinsics); their three backwards; the autograd add that merges the two gradients of w16
PATH_BYTES = {
    "slots_producer_fwd_bwd": 148 * 4,          # `slots * 1.0`: stand-in for the generator's last op on the [N,37] tensor
    "bone_stack": 60 * 2,                       # the 15 length columns -> [N,15]
    "fk_fwd": 148 + 60 + 12 + 192, "w2c_fwd": 192 + 192, "project_fwd": 192 + 36 + 128,
    "project_bwd": 192 + 36 + 128 + 192, "w2c_bwd": 192 + 192, "grad_add": 192 * 3, "fk_bwd": 148 + 60 + 12 + 192 + 148 + 12,
}


def _setup(n, dev, seed=0):
    import dhfk
    from dhfk import synthetic, tables
    import argparse
    d = synthetic.gan_like_torch(n, dev, seed=seed)
    slots = torch.zeros((n, 37), device=dev)
    slots[:, :33] = d["ang"]
    slots[:, 34:] = d["grot"]
    slots.requires_grad_(True)
    root = d["root"].clone().requires_grad_(True)
    bone_cols = [d["bone"][:, i].contiguous() for i in range(15)]
    blk = tables.camera_block("S1", 0)
    R = torch.tensor(blk[0:4], device=dev).view(1, 4)
    t = torch.tensor(blk[4:7], device=dev).view(1, 3)
    rows = torch.tensor(blk[7:16], device=dev).view(1, 9).repeat(n, 1)
    g = torch.Generator(device=dev).manual_seed(seed + 1)
    gw = torch.randn((n, 16, 3), generator=g, device=dev)
    gu = torch.randn((n, 16, 2), generator=g, device=dev)
    args = argparse.Namespace(batch_size=n, random_seed=0, single_or_multi_train_mode="single", architecture="3,3,3")
    fk = dhfk.Forward_Kinematics_DH_Model(args, ["S1"], None)
    idx = list(tables.H36M_32_To_16_Table)
    from dhfk import camera

    def step(accumulate=True):
        slots.grad = None          # what zero_grad() does: the gradients are assigned, not accumulated
        root.grad = None
        s = slots * 1.0            # the generator's tensor is a non-leaf
        kw = dict(right_leg_joints_angle=s[:, 0:5], left_leg_joints_angle=s[:, 5:10], body_joints_angle=s[:, 10:23],
                  right_hand_joints_angle=s[:, 23:28], left_hand_joints_angle=s[:, 28:33],
                  generator_global_rot_3d_pos_angle=s[:, -3:], root_3d_pos=root)
        for name, col in zip(BONES, bone_cols):
            kw[name] = col
        w32 = fk.change_3d_joint_angle(**kw)
        w16 = w32[:, idx]
        cam = camera.GAN_torch_world_to_camera(w16, R=R, t=t)
        uv = camera.project_to_2d(cam, rows)
        if accumulate:
            torch.autograd.backward((w16, uv), (gw, gu))
        else:                      # same graph, no AccumulateGrad nodes (they carry the stream the leaves were made on)
            return torch.autograd.grad((w16, uv), (slots, root), (gw, gu))
        return w16, uv

    return step


def measure(dev, big=1 << 20, small=(1024, 4608), peak_gbs=6528.7, profile=False):
    out = {"what": "change_3d_joint_angle -> [:, H36M_32_To_16_Table] -> GAN_torch_world_to_camera -> project_to_2d, "
                   "forward + backward through autograd (the unmodified caller's path)",
           "bytes_per_pose": sum(PATH_BYTES.values()), "bytes_breakdown": PATH_BYTES}
    step = _setup(big, dev)
    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 10
    e0.record()
    for _ in range(k):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / k
    gbs = out["bytes_per_pose"] * big / (ms * 1e-3) / 1e9
    out["at_%d" % big] = {"ms_per_step": ms, "poses_per_s": big / (ms * 1e-3), "hbm_gbs": gbs, "frac_of_copy_peak": gbs / peak_gbs}
    del step
    torch.cuda.empty_cache()
    for n in small:
        step = _setup(n, dev)
        for _ in range(30):
            step()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        reps = 300
        for _ in range(reps):
            step()
        torch.cuda.synchronize(dev)
        wall = (time.perf_counter() - t0) / reps
        e0.record()
        for _ in range(100):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
        rec = {"us_per_step_wall": wall * 1e6, "us_per_step_gpu_events": e0.elapsed_time(e1) / 100 * 1e3,
               "note": "wall = host-bound: three torch.autograd.Function nodes each way + the caller's own slicing"}
        # The same step captured once and replayed: what the GPU itself needs (the C-ABI calls are plain stream work).
        # First as it is; if the capture trips over the leaves' AccumulateGrad nodes ("legacy stream depends on a
        # capturing stream": they are bound to the stream the leaf tensors were created on), with the gradients taken
        # by torch.autograd.grad instead -- same kernels, same launches.
        for attempt, (mode, acc) in enumerate((("relaxed", True), ("relaxed", False), ("thread_local", False))):
          try:
            fn = (lambda: step(True)) if acc else (lambda: step(False))
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    keep = fn()
            torch.cuda.current_stream(dev).wait_stream(side)
            del keep
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode=mode):
                keep = fn()
            for _ in range(10):
                graph.replay()
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(200):
                graph.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            rec["us_per_step_cuda_graph"] = e0.elapsed_time(e1) / 200 * 1e3
            rec["cuda_graph_backward"] = ".backward() into .grad" if acc else "torch.autograd.grad (no AccumulateGrad nodes)"
            del graph, keep
            break
          except Exception as e:
            rec["us_per_step_cuda_graph"] = {"error": repr(e)[:200]}
            try:
                torch.cuda.synchronize(dev)
            except Exception:
                pass
        out["at_%d" % n] = rec
        if profile and n == small[0]:
            import cProfile
            import pstats
            import io
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(300):
                step()
            pr.disable()
            s = io.StringIO()
            pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
            out["profile_top"] = s.getvalue().splitlines()[:60]
    return out


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    peak = 6528.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    print(json.dumps(measure(dev, peak_gbs=peak, profile="--profile" in sys.argv), indent=1))
