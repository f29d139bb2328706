#!/usr/bin/env python
"""bench.py -- augmented poses/sec of the fused DH-FK + projection forward+backward path.

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch of synthetic poses: dhfk_forward
(world16 + uv16) followed by dhfk_backward (d angles, d global rotation, d root from upstream
gradients on world16 and uv16).

  --gpus 1   BASELINE.json configs[1]: 1,048,576 poses on one B200, S1/cam0, generator-range angles, template bone
             lengths x (1 +- 0.2), in-volume roots.
  --gpus N   BASELINE.json configs[4]: 16,777,216 poses per step sharded over the N ranks (strong scaling), and the
             GAN's gradient all-reduce -- generator + 3-D critic + 2-D critic (dense 256: 0.44 M + 0.88 M + 0.27 M
             parameters), ONE NCCL all-reduce on a persistent flat buffer (dhfk.parallel.FlatGradBuffer) -- INSIDE the
             timed step, issued on a side stream so that it overlaps the rank's FK kernels.  The same step at 1 M poses
             per rank (weak scaling) is reported next to it.  (--mode configs1 forces the N = 1 workload per rank.)

Rank 0 prints ONE JSON line (keys: see the task contract).  `value` is device-resident
throughput; `e2e` goes through the host-buffer C-ABI entry (pinned host memory in and out, copies
inside the timed region); `roofline` is for the dominant kernel (backward); `cpu_baseline` is the
torch port of the reference's own CPU path timed on this host.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

# stdout must carry exactly ONE JSON line.  Libraries (NCCL prints its version banner, torchrun children
# inherit the pipe) write to fd 1 as well, so fd 1 is pointed at stderr for the whole run and the JSON
# line goes to a private duplicate of the original stdout.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit_json(obj):
    os.write(_JSON_FD, (json.dumps(obj) + "\n").encode())


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(device_index):
    """Pin this rank's host threads to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned buffer is
    allocated (first touch places the pages there).  The e2e leg is a PCIe/host-memory stream: with several ranks on
    one box, buffers on the far socket halve it.  Returns a short description (or None when sysfs says nothing)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev_id = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev_id)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception:
        return None


METRIC = "augmented poses/sec (FK+proj fwd+bwd)"
UNIT = "poses/s"
FWD_BYTES = 536    # per pose: 54 floats in, 48 + 32 floats out           (SURVEY 8d / BASELINE.md 4)
BWD_BYTES = 692    # per pose: 54 + 48 + 32 floats in, 33 + 3 + 3 floats out
HBM_FALLBACK_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--poses", type=int, default=1 << 20, help="poses per GPU per step")
    ap.add_argument("--fast-trig", action="store_true", help="MUFU sin/cos in the forward too (DHFK_FLAG_FAST_TRIG)")
    ap.add_argument("--accurate-trig", action="store_true", help="table sincos in the backward too (DHFK_FLAG_ACCURATE_TRIG)")
    ap.add_argument("--buffers", type=int, default=4, help="rotating input buffer sets (L2 hygiene)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-chunk", type=int, default=1024, help="reference arm: poses per torch call")
    ap.add_argument("--ref-chunks", type=int, default=4, help="reference arm: chunks per step")
    ap.add_argument("--mode", default="auto", choices=["auto", "configs1", "configs4"],
                    help="auto: configs[1] at --gpus 1, configs[4] (16M poses sharded + in-step gradient all-reduce) otherwise")
    ap.add_argument("--total-poses", type=int, default=1 << 24, help="configs[4]: poses per step over all ranks")
    ap.add_argument("--no-extras", action="store_true", help="skip side-kernel / drop-in path / GAN-step extras")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 17, help="rows per chunk of the host-buffer pipeline")
    ap.add_argument("--e2e-slots", type=int, default=3, help="device slots of the host-buffer pipeline")
    ap.add_argument("--nccl-max-ctas", type=int, default=int(os.environ.get("DHFK_NCCL_MAX_CTAS", "0")),
                    help="CTAs of a dedicated NCCL communicator for the gradient all-reduce; 0 (default) = NCCL's default "
                         "communicator, which measured best (profiles/r2k_nccl_sweep.txt: fewer CTAs make the 6 MB "
                         "all-reduce 2-6x slower, too slow to hide behind a 0.2 ms step)")
    ap.add_argument("--exchange", default=os.environ.get("DHFK_EXCHANGE", "peer"), choices=["peer", "nccl"],
                    help="gradient exchange at --gpus > 1: peer = dhfk_grad_allreduce (one hand-written kernel over NVLink "
                         "peer memory / NVLS multicast; falls back to NCCL where symmetric memory is unavailable), nccl = "
                         "ncclAllReduce")
    ap.add_argument("--exchange-ctas", type=int, default=int(os.environ.get("DHFK_EXCHANGE_CTAS", "16")),
                    help="CTAs of dhfk_grad_allreduce")
    ap.add_argument("--exchange-threads", type=int, default=int(os.environ.get("DHFK_EXCHANGE_THREADS", "512")),
                    help="threads per CTA of dhfk_grad_allreduce")
    ap.add_argument("--exchange-priority", type=int, default=int(os.environ.get("DHFK_EXCHANGE_PRIORITY", "-1")),
                    help="priority of the side stream the exchange runs on (-1 = high)")
    return ap.parse_args()


def resolve_mode(args):
    if args.mode == "auto":
        return "configs1" if args.gpus <= 1 else "configs4"
    return args.mode


def workload_config(args, extra=None):
    strong = resolve_mode(args) == "configs4"
    cfg = {
        "workload": ("BASELINE configs[4]: data-parallel GAN augmentation sweep, %d poses per step sharded over %d B200 "
                     "(fused DH-FK + projection forward+backward per shard) with the NCCL gradient all-reduce of generator + "
                     "3-D critic + 2-D critic (dense 256) inside the step" % (args.total_poses, args.gpus)) if strong else
                    ("BASELINE configs[1]: fused DH-FK + projection forward+backward, 1M poses per B200, "
                     "16-joint H36M skeleton, synthetic generator-range angles + template bone lengths, S1/cam0"),
        "poses_per_gpu": (args.total_poses // max(args.gpus, 1)) if strong else args.poses,
        "outputs": "world16[N,16,3]+uv16[N,16,2]; grads d_ang[N,33]+d_grot[N,3]+d_root[N,3]",
        "bytes_per_pose": {"forward": FWD_BYTES, "backward": BWD_BYTES},
        "parallelism": ("dp%d (rows sharded with dhfk.parallel.shard_rows; one NCCL all-reduce of the flat gradient buffer "
                        "per step, overlapped with the FK kernels on a side stream)" % args.gpus) if strong else
                       ("dp%d (rows sharded, no data-path collective)" % args.gpus),
        "trig": ("fwd+bwd MUFU.SIN/COS (DHFK_FLAG_FAST_TRIG)" if args.fast_trig else
                 "fwd polynomial + bwd table sincos, <=1e-7 (DHFK_FLAG_ACCURATE_TRIG)" if getattr(args, "accurate_trig", False) else
                 "library default: fwd polynomial (abs err 7.7e-8), bwd MUFU.SIN/COS after exact degree reduction "
                 "(abs err 4e-7; gradients 1.8e-7 from exact at 1M poses, tolerance 1e-5)"),
    }
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls NVML for SM clock / throttle reasons while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {
                "nvmlClocksEventReasonSwPowerCap": "sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap": "sw_power_cap",
                "nvmlClocksEventReasonHwSlowdown": "hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown": "hw_slowdown",
                "nvmlClocksEventReasonSwThermalSlowdown": "sw_thermal_slowdown",
                "nvmlClocksThrottleReasonSwThermalSlowdown": "sw_thermal_slowdown",
                "nvmlClocksEventReasonHwThermalSlowdown": "hw_thermal_slowdown",
                "nvmlClocksThrottleReasonHwThermalSlowdown": "hw_thermal_slowdown",
                "nvmlClocksEventReasonHwPowerBrakeSlowdown": "hw_power_brake",
                "nvmlClocksThrottleReasonHwPowerBrakeSlowdown": "hw_power_brake",
            }
            masks = {}
            for attr, nm in names.items():
                if hasattr(pynvml, attr):
                    masks[int(getattr(pynvml, attr))] = nm
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop_evt.is_set():
                self.samples.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for m, nm in masks.items():
                    if r & m:
                        self.reasons.add(nm)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report it rather than fail the bench
            self.err = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(2.0)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), **({"error": self.err} if self.err else {})}


def physical_gpu_index(local_index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


# ---------------------------------------------------------------------------------------------------
def time_torch_port(chunk, chunks, steps, warmup, threads=None):
    """Reference arm / cpu_baseline: the torch port of the reference's own CPU path (oracle/torch_port.py,
    bit-identical to the reference on the build host), forward + backward, `chunks` calls of `chunk` poses
    per step.  Returns (poses_per_s, seconds_per_step, threads)."""
    import numpy as np
    import torch
    import torch_port
    from dhfk import synthetic, tables
    torch.set_num_threads(threads or host_threads())
    threads = torch.get_num_threads()
    blk = tables.camera_block("S1", 0)
    inp = synthetic.gan_like(chunk, seed=1234)
    up = synthetic.upstream_grads(chunk, seed=4321)
    ang, grot, bone, root = (torch.tensor(inp[k]) for k in ("ang", "grot", "bone", "root"))
    gw, gu = torch.tensor(up["g_world"]), torch.tensor(up["g_uv"])

    def one_step():
        for _ in range(chunks):
            a = ang.clone().requires_grad_(True); g = grot.clone().requires_grad_(True)
            r = root.clone().requires_grad_(True)
            _, w16, _, uv = torch_port.pipeline(a, g, bone, r, blk)
            ((w16 * gw).sum() + (uv * gu).sum()).backward()

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return chunk * chunks / dt, dt, threads


def time_torch_port_forward_only(chunk=1024, reps=5, warmup=2, threads=None):
    """SURVEY 8d asks for the reference's forward (no_grad) next to forward+backward: best of `reps` calls."""
    import torch
    import torch_port
    from dhfk import synthetic, tables
    torch.set_num_threads(threads or host_threads())
    blk = tables.camera_block("S1", 0)
    inp = synthetic.gan_like(chunk, seed=1234)
    ang, grot, bone, root = (torch.tensor(inp[k]) for k in ("ang", "grot", "bone", "root"))
    best = float("inf")
    with torch.no_grad():
        for it in range(warmup + reps):
            t0 = time.perf_counter()
            torch_port.pipeline(ang, grot, bone, root, blk)
            if it >= warmup:
                best = min(best, time.perf_counter() - t0)
    return chunk / best, torch.get_num_threads()


def time_torch_port_gpu(dev, chunk=16384, reps=3):
    """Context only: the reference's op sequence (oracle/torch_port.py) in torch EAGER on the GPU, every tensor created
    on the device (kinder than the reference's own CUDA branch, which builds each of its 34 matrices on the host and
    copies it over).  ~7 000 ATen launches forward, ~13 000 with autograd: launch-bound at any batch that fits."""
    import torch
    import torch_port
    from dhfk import synthetic, tables
    blk = tables.camera_block("S1", 0)
    inp = synthetic.gan_like(chunk, seed=1234)
    up = synthetic.upstream_grads(chunk, seed=4321)
    ang, grot, bone, root = (torch.tensor(inp[k], device=dev) for k in ("ang", "grot", "bone", "root"))
    gw, gu = torch.tensor(up["g_world"], device=dev), torch.tensor(up["g_uv"], device=dev)
    best = float("inf")
    for it in range(reps + 1):
        a = ang.clone().requires_grad_(True); g = grot.clone().requires_grad_(True); r = root.clone().requires_grad_(True)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        _, w16, _, uv = torch_port.pipeline(a, g, bone, r, blk)
        ((w16 * gw).sum() + (uv * gu).sum()).backward()
        torch.cuda.synchronize(dev)
        if it > 0:
            best = min(best, time.perf_counter() - t0)
    return chunk / best, best


def time_c_oracle(n=131072):
    """Extra context: the float64 C oracle (OpenMP, all cores) forward+backward."""
    import c_oracle
    from dhfk import synthetic, tables
    c_oracle.set_num_threads(host_threads())
    inp = synthetic.gan_like(n, seed=1)
    up = synthetic.upstream_grads(n, seed=2)
    blk = tables.camera_block("S1", 0)
    c_oracle.forward(inp["ang"][:1024], inp["grot"][:1024], inp["bone"][:1024], inp["root"][:1024], blk)
    t0 = time.perf_counter()
    c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk)
    c_oracle.backward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk, g_world=up["g_world"], g_uv=up["g_uv"],
                      want_bone=False)
    dt = time.perf_counter() - t0
    return n / dt, c_oracle.num_threads()


def time_reference_pipeline(chunk, chunks, steps, warmup, threads=None):
    """The reference's OWN code -- Forward_Kinematics_DH_Model.change_3d_joint_angle -> [:, H36M_32_To_16_Table] ->
    GAN_torch_world_to_camera -> project_to_2d and autograd through them -- imported unmodified from the archive
    oracle/stage_ref.py staged (oracle/_ref/dh_aug_ref.zip), on the host cores.  The model object is built once, as
    run_Fk_GAN.py does.  Returns (poses_per_s, seconds_per_step, threads) or None when the archive is absent."""
    import torch
    import ref_harness as rh
    if rh.reference_source(prefer_staged=True) is None:
        return None
    from dhfk import synthetic, tables
    torch.set_num_threads(threads or host_threads())
    threads = torch.get_num_threads()
    ref = rh.import_reference(force_cpu=True, prefer_staged=True)
    blk = tables.camera_block("S1", 0)
    inp = synthetic.gan_like(chunk, seed=1234)
    up = synthetic.upstream_grads(chunk, seed=4321)
    ang, grot, bone, root = (torch.tensor(inp[k]) for k in ("ang", "grot", "bone", "root"))
    gw, gu = torch.tensor(up["g_world"]), torch.tensor(up["g_uv"])
    model = ref.fk.Forward_Kinematics_DH_Model(rh.make_args(chunk), ["S1"], None)
    q, t = torch.tensor(blk[0:4]).view(1, 4), torch.tensor(blk[4:7]).view(1, 3)
    rows = torch.tensor(blk[7:16]).view(1, 9).repeat(chunk, 1)
    idx = ref.h36m.H36M_32_To_16_Table

    def one_step():
        for _ in range(chunks):
            a = ang.clone().requires_grad_(True); g = grot.clone().requires_grad_(True)
            r = root.clone().requires_grad_(True)
            kw = dict(right_leg_joints_angle=a[:, 0:5], left_leg_joints_angle=a[:, 5:10], body_joints_angle=a[:, 10:23],
                      right_hand_joints_angle=a[:, 23:28], left_hand_joints_angle=a[:, 28:33],
                      generator_global_rot_3d_pos_angle=g, root_3d_pos=r)
            for i, name in enumerate(rh.BONE_KWARGS):
                kw[name] = bone[:, i]
            w16 = model.change_3d_joint_angle(**kw)[:, idx]
            uv = ref.camera.project_to_2d(ref.camera.GAN_torch_world_to_camera(w16, R=q, t=t), rows)
            ((w16 * gw).sum() + (uv * gu).sum()).backward()

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return chunk * chunks / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    # bound the run: a step is ref_chunks x ref_chunk poses (~0.1 s per 1024-pose call on 8 cores)
    kind, what = "reference", ("the UNMODIFIED reference (oracle/_ref/dh_aug_ref.zip: Forward_Kinematics_DH_Model."
                               "change_3d_joint_angle -> 32->16 gather -> GAN_torch_world_to_camera -> project_to_2d, "
                               "torch autograd), host cores only")
    res = time_reference_pipeline(args.ref_chunk, args.ref_chunks, steps, warmup)
    if res is None:
        kind, what = "port", ("oracle/torch_port.py: the reference's own torch op sequence (bit-identical to /root/reference "
                              "on the build host; the staged archive is absent), host cores only")
        res = time_torch_port(args.ref_chunk, args.ref_chunks, steps, warmup)
    pps, dt, threads = res
    sample = "%d x %d-pose torch calls per step (reference batch size), fwd+bwd, CPU" % (args.ref_chunks, args.ref_chunk)
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if resolve_mode(args) == "configs4" else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, {"reference_arm": what}),
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# ---------------------------------------------------------------------------------------------------
class HotPath:
    """Device buffers + the two C-ABI calls of one step over `n` poses (rotating over `nbuf` input sets)."""

    def __init__(self, lib, n, dev, nbuf, seed, flags, cam_ptr, stream_ptr):
        import torch
        from dhfk import synthetic
        self.lib, self.n, self.flags, self.cam_ptr, self.sp, self.nbuf = lib, n, flags, cam_ptr, stream_ptr, nbuf
        self.sets = []
        for b in range(nbuf):
            d = synthetic.gan_like_torch(n, dev, seed=seed + b)
            g = torch.Generator(device=dev).manual_seed(seed + 3087 + b)
            d["g_world"] = torch.randn((n, 16, 3), generator=g, device=dev)
            d["g_uv"] = torch.randn((n, 16, 2), generator=g, device=dev)
            self.sets.append(d)
        self.world = torch.empty((n, 16, 3), device=dev); self.uv = torch.empty((n, 16, 2), device=dev)
        self.g_ang = torch.empty((n, 33), device=dev); self.g_grot = torch.empty((n, 3), device=dev)
        self.g_root = torch.empty((n, 3), device=dev)

    def fwd(self, i, flags=None):
        from dhfk import _cabi
        d = self.sets[i % self.nbuf]
        rc = self.lib.dhfk_forward(d["ang"].data_ptr(), 33, d["grot"].data_ptr(), 3, d["bone"].data_ptr(), 15,
                                   d["root"].data_ptr(), 3, self.cam_ptr, self.world.data_ptr(), None,
                                   self.uv.data_ptr(), self.n, self.flags if flags is None else flags, self.sp)
        _cabi.check(rc, "dhfk_forward")

    def bwd(self, i, flags=None):
        from dhfk import _cabi
        d = self.sets[i % self.nbuf]
        rc = self.lib.dhfk_backward(d["ang"].data_ptr(), 33, d["grot"].data_ptr(), 3, d["bone"].data_ptr(), 15,
                                    d["root"].data_ptr(), 3, self.cam_ptr, d["g_world"].data_ptr(), None,
                                    d["g_uv"].data_ptr(), self.g_ang.data_ptr(), 33, self.g_grot.data_ptr(), 3,
                                    self.g_root.data_ptr(), 3, None, 15, self.n, self.flags if flags is None else flags, self.sp)
        _cabi.check(rc, "dhfk_backward")

    def step(self, i, flags=None):
        self.fwd(i, flags); self.bwd(i, flags)


def gan_models(dev):
    """Generator + 3-D critic + 2-D critic at the sizes of SURVEY 5 (dense 256): what a data-parallel GAN step all-reduces."""
    import argparse
    import torch
    from dhfk import Fk_discriminator, Fk_generator
    a = argparse.Namespace(batch_size=1024, GAN_OUTPUT_DIM=35, Gen_DenseDim=256, Dis_DenseDim_3D=256, Dis_DenseDim_2D=256,
                           GAN_whether_use_preAngle=True, whether_use_RT=True, bone_len_scaler="different")
    torch.manual_seed(0)
    G = Fk_generator.Fk_Generator(None, a, dev).to(dev)
    D3 = Fk_discriminator.Fk_3D_Discriminator(dev, a).to(dev)
    D2 = Fk_discriminator.Fk_2D_Discriminator(a).to(dev)
    return G, D3, D2


def timed_steps(path, steps, stream, dev, barrier, allreduce=None, probe_every=0):
    """K steps bracketed by CUDA events on the launching stream.  allreduce: (FlatGradBuffer, side stream) -> one
    collective per step on the side stream, started when the step starts and joined before the step ends."""
    import torch
    probes = {}
    if probe_every:
        probes = {i: [torch.cuda.Event(enable_timing=True) for _ in range(3)] for i in range(0, steps, probe_every)}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    go = torch.cuda.Event()
    barrier()
    ev0.record(stream)
    for i in range(steps):
        if allreduce is not None:
            buf, side = allreduce
            go.record(stream)
            side.wait_event(go)
            with torch.cuda.stream(side):
                buf.allreduce()
        pr = probes.get(i)
        if pr is None:
            path.step(i)
        else:
            pr[0].record(stream); path.fwd(i)
            pr[1].record(stream); path.bwd(i)
            pr[2].record(stream)
        if allreduce is not None:
            stream.wait_stream(allreduce[1])
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    total_ms = ev0.elapsed_time(ev1)
    fwd_ms = bwd_ms = None
    if probes:
        fwd_ms = sum(e[0].elapsed_time(e[1]) for e in probes.values()) / len(probes)
        bwd_ms = sum(e[1].elapsed_time(e[2]) for e in probes.values()) / len(probes)
    return total_ms, fwd_ms, bwd_ms


def feed_sampler(sampler, path, dev):
    """Keep the clock sampler fed when the timed region was shorter than a few polling periods: the same step (no
    collective, so the ranks need not agree on a count) for one more second on the sampled rank.  True if it ran."""
    import torch
    if sampler is None or len(sampler.samples) >= 10:
        return False
    t_end = time.time() + 1.0
    i = 0
    while time.time() < t_end:
        path.step(i); i += 1
        if i % 32 == 0:
            torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    return True


EXPOSED_HOW = ("after 0.3 s of the same steps (the board's burst regime is over by then), five K-step runs without / with / "
               "without / with / without the exchange: mean of the two with it - mean of the three without it, per step")


def exchange_exposure(path, steps, stream, dev, barrier, ar, distributed):
    """(exposed ms per step, ms per step without the exchange) in the board's steady state.  Measured apart from the
    headline run: a run out of idle sits in the burst regime and the clocks step down ~0.1 s into a series, which a
    with / without pair straddling that moment reads as +-5 % (profiles/r2t3_bench_n2.json).  Symmetric order: a linear
    drift cancels."""
    import torch
    t_end = time.time() + 0.3
    i = 0
    while time.time() < t_end:
        path.step(i); i += 1
        if i % 16 == 0:
            torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    runs = []
    for k in range(5):
        ms, _, _ = timed_steps(path, steps, stream, dev, barrier, allreduce=(ar if k % 2 else None))
        runs.append(ms)
    runs = max_over_ranks(runs, dev, distributed)
    plain = (runs[0] + runs[2] + runs[4]) / 3 / steps
    return (runs[1] + runs[3]) / 2 / steps - plain, plain


def max_over_ranks(vals, dev, distributed):
    import torch
    import torch.distributed as dist
    t = torch.tensor([v if v is not None else -1.0 for v in vals], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) if x >= 0 else None for x in t.tolist()]


def copy_ceiling(n, dev, chunk, reps=10, split=False):
    """A bare pinned-memory H2D + D2H of the bytes the e2e step moves (536 B/row up, 476 B/row down), in the same chunk
    sizes, on two streams with no kernel and no dependency between them: the PCIe / host-memory floor of `e2e`.
    split=True moves them as the e2e entry does -- six arrays up (33, 3, 15, 3, 48, 32 floats per row), five down
    (48, 32, 33, 3, 3) -- instead of one buffer each way."""
    import torch
    ups, downs = ((33, 3, 15, 3, 48, 32), (48, 32, 33, 3, 3)) if split else ((134,), (119,))
    h_up = [torch.empty(n * w, dtype=torch.float32).pin_memory() for w in ups]
    h_dn = [torch.empty(n * w, dtype=torch.float32).pin_memory() for w in downs]
    d_up = [torch.empty(chunk * w, dtype=torch.float32, device=dev) for w in ups]
    d_dn = [torch.empty(chunk * w, dtype=torch.float32, device=dev) for w in downs]
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def once():
        for r0 in range(0, n, chunk):
            rows = min(chunk, n - r0)
            with torch.cuda.stream(s_up):
                for w, h, d in zip(ups, h_up, d_up):
                    d[:rows * w].copy_(h[r0 * w:(r0 + rows) * w], non_blocking=True)
            with torch.cuda.stream(s_dn):
                for w, h, d in zip(downs, h_dn, d_dn):
                    h[r0 * w:(r0 + rows) * w].copy_(d[:rows * w], non_blocking=True)
        s_up.synchronize(); s_dn.synchronize()

    once()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    return (time.perf_counter() - t0) / reps


def reference_loop_timings(timeout=420):
    """BASELINE configs[2] / configs[3] through the reference's OWN loops (oracle/ref_loop.py, test infrastructure, run in
    child processes): unpatched on this GPU where its CUDA branch runs at all, then with dhfk.dropin.install(all)."""
    import subprocess
    out = {}
    script = os.path.join(ROOT, "oracle", "ref_loop.py")
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    for key, inst in (("reference_unpatched_same_gpu", "none"), ("reference_loop_with_dropin", "all")):
        try:
            r = subprocess.run([sys.executable, script, "--device", "cuda", "--install", inst, "--iters", "10"],
                               capture_output=True, text=True, timeout=timeout, env=env)
            lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
            out[key] = json.loads(lines[-1]) if lines else {"error": "rc=%d %s" % (r.returncode, r.stderr[-300:])}
        except Exception as e:
            out[key] = {"error": repr(e)[:300]}
    return out


def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dhfk
    from dhfk import _cabi, parallel, synthetic, tables

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world_size > 1
    numa = bind_to_gpu_numa_node(local_rank) if distributed else None
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    strong = resolve_mode(args) == "configs4"

    lib = _cabi.load()
    flags = (_cabi.FLAG_FAST_TRIG if args.fast_trig else 0) | (_cabi.FLAG_ACCURATE_TRIG if args.accurate_trig else 0)
    blk = tables.camera_block("S1", 0)
    cam_ptr = blk.ctypes.data
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    nbuf = max(1, args.buffers)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the 1M-poses-per-GPU workload (configs[1]; the weak-scaling companion of configs[4])
    n = args.poses
    path1 = HotPath(lib, n, dev, nbuf, 1234 + 97 * rank, flags, cam_ptr, sp)
    fwd, bwd = path1.fwd, path1.bwd
    sets, world, uv, g_ang, g_grot, g_root = path1.sets, path1.world, path1.uv, path1.g_ang, path1.g_grot, path1.g_root

    # parity before any timing is reported: the first rows of buffer set 0 against the float64 C oracle,
    # as the checker only (part of the cpu_baseline leg: skipped with --no-cpu-baseline)
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        import c_oracle
        m = min(n, 4096)
        d = sets[0]
        fwd(0); bwd(0)
        torch.cuda.synchronize(dev)
        host = {k: d[k][:m].cpu().numpy() for k in ("ang", "grot", "bone", "root", "g_world", "g_uv")}
        o = c_oracle.forward(host["ang"], host["grot"], host["bone"], host["root"], blk)
        b = c_oracle.backward(host["ang"], host["grot"], host["bone"], host["root"], blk, g_world=host["g_world"],
                              g_uv=host["g_uv"], want_bone=False)
        rel = lambda x, r: float((np.abs(x[:m].cpu().numpy().astype(np.float64) - r) / np.maximum(np.abs(r), 1.0)).max())
        errs = {"world": rel(world, o["world16"]), "uv": rel(uv, o["uv"]), "g_ang": rel(g_ang, b["g_ang"]),
                "g_grot": rel(g_grot, b["g_grot"]), "g_root": rel(g_root, b["g_root"])}
        if not all(v <= 1e-5 for v in errs.values()):
            raise SystemExit("bench.py: parity check against the oracle failed before timing: %r" % errs)
        parity = {"checked_poses": m, "tolerance": 1e-5, "max_rel_err": errs, "oracle": "oracle/dhfk_oracle.c (float64)"}

    steps, warmup = max(args.steps, 1), max(args.warmup, 3)

    # the gradient exchange of the data-parallel GAN step (N > 1): generator + both critics in ONE flat buffer
    gbuf = side = None
    allreduce_extra = None
    if distributed:
        G, D3, D2 = gan_models(dev)
        gparams = [*G.parameters(), *D3.parameters(), *D2.parameters()]
        gbuf = parallel.FlatGradBuffer(gparams, peer_exchange=(args.exchange == "peer"), max_ctas=args.exchange_ctas,
                                       cta_threads=args.exchange_threads)
        side = torch.cuda.Stream(dev, priority=args.exchange_priority)   # -1: the exchange's CTAs go ahead of the FK tiles queued beside them
        ggroup = parallel.grad_allreduce_group(args.nccl_max_ctas) if gbuf.peer is None else None
        gbuf_allreduce = gbuf.allreduce
        gbuf.allreduce = lambda: gbuf_allreduce(group=ggroup)        # every call below goes through the chosen path

        def time_alone(fn, reps=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize(dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a0.record(stream)
            for _ in range(reps):
                fn()
            a1.record(stream)
            torch.cuda.synchronize(dev)
            return max_over_ranks([a0.elapsed_time(a1) / reps], dev, True)[0]

        nccl_buf = torch.empty_like(gbuf.flat)
        exchange_check = None
        if gbuf.peer is not None:      # the hand-written exchange against NCCL on this run's own buffer, before timing
            gen = torch.Generator(device=dev).manual_seed(4242 + rank)
            gbuf.flat.copy_(torch.randn(gbuf.flat.numel(), generator=gen, device=dev))
            nccl_buf.copy_(gbuf.flat)
            torch.distributed.all_reduce(nccl_buf, op=torch.distributed.ReduceOp.AVG)
            gbuf.allreduce()
            torch.cuda.synchronize(dev)
            gbuf.peer.check()
            err = float((gbuf.flat - nccl_buf).abs().max())
            first = gbuf.flat.clone()
            torch.distributed.broadcast(first, src=0)
            same = bool(torch.equal(first, gbuf.flat))
            (err,) = max_over_ranks([err], dev, True)
            (diff,) = max_over_ranks([0.0 if same else 1.0], dev, True)
            if err > 1e-5 or diff:
                raise SystemExit("bench.py: dhfk_grad_allreduce differs from NCCL (max abs %g, ranks identical: %s)" % (err, not diff))
            exchange_check = {"max_abs_err_vs_nccl": err, "bit_identical_across_ranks": True}
        gbuf.flat.fill_(1.0)
        nel = gbuf.flat.numel()
        alone_ms = time_alone(gbuf.allreduce)
        nccl_ms = time_alone(lambda: torch.distributed.all_reduce(nccl_buf, op=torch.distributed.ReduceOp.AVG)) \
            if gbuf.peer is not None else alone_ms
        if gbuf.peer is not None:
            comm = "dhfk_grad_allreduce: one kernel over NVLink peer memory (%s), <= %d CTAs" % (
                "NVLS multimem.ld_reduce / multimem.st" if gbuf.peer.multicast else "peer loads / stores", gbuf.peer.max_ctas) + \
                " of %d threads, side stream priority %d" % (gbuf.peer.cta_threads, args.exchange_priority)
            what = ("dhfk.parallel.FlatGradBuffer.allreduce: generator + 3-D critic + 2-D critic gradients (dense 256) live "
                    "in one persistent symmetric buffer (the slices are the .grad tensors); every rank reduces its slice "
                    "in place through the NVSwitch and writes it to all ranks, two in-kernel cross-GPU barriers, no NCCL "
                    "call; ms_alone = back to back on an idle GPU, ms_alone_nccl = ncclAllReduce(AVG) on the same bytes")
        else:
            comm = ("NCCL, dedicated communicator, max_ctas=%d" % args.nccl_max_ctas) if ggroup is not None else "NCCL, default communicator"
            what = ("dhfk.parallel.FlatGradBuffer.allreduce: generator + 3-D critic + 2-D critic gradients (dense 256) live in "
                    "one persistent buffer (the slices are the .grad tensors), one ncclAllReduce(AVG), no pack / unpack "
                    "kernels; ms_alone = back to back on an idle GPU")
            if args.exchange == "peer":
                what += "; peer exchange unavailable here: %s" % (parallel.PeerExchange.last_error,)
        allreduce_extra = {"bytes": int(nel) * 4, "ms_alone": alone_ms, "ms_alone_nccl": nccl_ms, "backend": "nccl",
                           "collectives_per_step": 1, "communicator": comm, "what": what}
        if exchange_check:
            allreduce_extra["parity"] = exchange_check

    sampler = ClockSampler(physical_gpu_index(local_rank)) if rank == 0 else None
    probe_every = 8 if steps >= 16 else 1
    extension = False
    weak_extra = None
    if strong:
        lo, hi = parallel.shard_rows(args.total_poses, rank, world_size)
        n_big = hi - lo
        path = HotPath(lib, n_big, dev, 1 if n_big >= (1 << 21) else 2, 777 + 97 * rank, flags, cam_ptr, sp)
        ar = (gbuf, side) if distributed else None
        # the headline first, out of idle like the N = 1 line: K steps with the exchange inside them
        for i in range(warmup):
            path.step(i)
        if sampler:
            sampler.start()
        total_ms, fwd_ms, bwd_ms = timed_steps(path, steps, stream, dev, barrier, allreduce=ar, probe_every=probe_every)
        barrier()
        extension = feed_sampler(sampler, path, dev)
        if sampler:
            sampler.stop()
        total_ms, fwd_ms, bwd_ms = max_over_ranks([total_ms, fwd_ms, bwd_ms], dev, distributed)
        n_step = n_big
        value = args.total_poses / (total_ms / steps * 1e-3)
        if allreduce_extra is not None:
            exposed, plain = exchange_exposure(path, steps, stream, dev, barrier, ar, distributed)
            allreduce_extra["ms_exposed_per_step"] = exposed
            allreduce_extra["ms_per_step_without_allreduce"] = plain
            allreduce_extra["exposed_how"] = EXPOSED_HOW
        # weak companion (1 M poses per rank, the same in-step exchange): its value out of idle as well, then its exposure
        torch.cuda.synchronize(dev)
        time.sleep(0.5)
        for i in range(warmup):
            path1.step(i)
        w_ms, _, _ = timed_steps(path1, steps, stream, dev, barrier, allreduce=ar)
        (w_ms,) = max_over_ranks([w_ms], dev, distributed)
        weak_extra = {"poses_per_gpu": n, "value": n * world_size / (w_ms / steps * 1e-3), "unit": UNIT,
                      "ms_per_step": w_ms / steps,
                      "what": "weak scaling: 1,048,576 poses per rank + the same in-step gradient exchange, %d steps after "
                              "0.5 s of idle and the warm-ups" % steps}
        if ar is not None:
            exposed, plain = exchange_exposure(path1, steps, stream, dev, barrier, ar, distributed)
            weak_extra["ms_exposed_per_step"] = exposed
            weak_extra["ms_per_step_without_allreduce"] = plain
            if gbuf.peer is not None:
                gbuf.peer.check()          # no exchange in the timed regions gave up on a peer
    else:
        path = path1
        for i in range(warmup):
            path.step(i)
        barrier()
        if sampler:
            sampler.start()
        # One event pair brackets the K timed steps; every `probe_every`-th step additionally carries three events around
        # its two launches for the live per-kernel durations (an event record is a ~2 us bubble on the stream, so probing
        # every launch would cost the headline ~3 %).
        total_ms, fwd_ms, bwd_ms = timed_steps(path, steps, stream, dev, barrier, probe_every=probe_every)
        barrier()
        extension = feed_sampler(sampler, path, dev)
        if sampler:
            sampler.stop()
        total_ms, fwd_ms, bwd_ms = max_over_ranks([total_ms, fwd_ms, bwd_ms], dev, distributed)
        n_step = n
        value = n * world_size / (total_ms / steps * 1e-3)
        if distributed:      # forced configs1 at N > 1: report the collective next to the collective-free step
            ar_ms, _, _ = timed_steps(path, steps, stream, dev, barrier, allreduce=(gbuf, side))
            (ar_ms,) = max_over_ranks([ar_ms], dev, True)
            allreduce_extra["ms_exposed_per_step"] = (ar_ms - total_ms) / steps
            if gbuf.peer is not None:
                gbuf.peer.check()
    ms_per_step = total_ms / steps
    del path
    if strong:
        torch.cuda.empty_cache()

    # ---- e2e through the host-buffer C-ABI entry (pinned host in/out, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        hin = synthetic.gan_like(min(n, 1 << 16), seed=5 + rank)      # tile a 64k-pose draw to N (host RNG is slow)
        rep = (n + hin["ang"].shape[0] - 1) // hin["ang"].shape[0]
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:n])).pin_memory()
        h_ang, h_grot, h_bone, h_root = (pin(hin[k]) for k in ("ang", "grot", "bone", "root"))
        up = synthetic.upstream_grads(min(n, 1 << 16), seed=6 + rank)
        h_gw, h_gu = pin(up["g_world"]), pin(up["g_uv"])
        out = {}
        e2e_steps = max(3, min(steps, 10))
        chunk, slots = args.e2e_chunk, args.e2e_slots
        call = lambda o, gw_, gu_: dhfk.fk_project_host(h_ang, h_grot, h_bone, h_root, blk, gw_, gu_, chunk_rows=chunk,
                                                        num_streams=slots, workspace=o.get("_workspace"), out=o,
                                                        fast_trig=args.fast_trig, accurate_grad=args.accurate_trig)
        for _ in range(2):
            out = call(out, h_gw, h_gu)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = call(out, h_gw, h_gu)
        torch.cuda.synchronize(dev)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        # forward only (what the reference itself moves: only the fake pairs ever leave the GPU, model_fk_gan_train.py:486-488)
        fo = {}
        for _ in range(2):
            fo = call(fo, None, None)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fo = call(fo, None, None)
        torch.cuda.synchronize(dev)
        fo_s = (time.perf_counter() - t0) / e2e_steps
        barrier()
        ceil_s = copy_ceiling(n, dev, chunk)
        barrier()
        ceil_split_s = copy_ceiling(n, dev, chunk, split=True)
        e2e_s, fo_s, ceil_s, ceil_split_s = max_over_ranks([e2e_s, fo_s, ceil_s, ceil_split_s], dev, distributed)
        h2d, d2h = n * 536, n * 476
        e2e = {"value": n * world_size / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "poses_per_rank": n,
               "copy_ceiling_ms": ceil_s * 1e3, "frac_of_ceiling": ceil_s / e2e_s,
               "copy_ceiling_same_arrays_ms": ceil_split_s * 1e3,     # the same bytes as the entry's 6 + 5 separate arrays
               "per_rank_gbs": {"h2d": h2d / e2e_s / 1e9, "d2h": d2h / e2e_s / 1e9,
                                "ceiling_h2d": h2d / ceil_s / 1e9, "ceiling_d2h": d2h / ceil_s / 1e9},
               "copy_ceiling": "bare pinned H2D (536 B/row) + D2H (476 B/row) of the same bytes in the same %d-row chunks on "
                               "two streams, no kernels, same ranks active: the PCIe / host-memory floor" % chunk,
               "forward_only": {"value": n * world_size / fo_s, "unit": UNIT, "ms_per_step": fo_s * 1e3,
                                "h2d_bytes_per_step": n * 216, "d2h_bytes_per_step": n * 320,
                                "what": "same entry with no upstream gradients: inputs up, world16 + uv16 down"},
               "api": "dhfk.fk_project_host -> dhfk_forward_backward_host (pinned host buffers, %d-row chunks through an "
                      "upload / compute / download stream pipeline over %d device slots)" % (chunk, slots)}
        if strong:
            e2e["workload_note"] = "measured on 1,048,576 poses per rank (the configs[1] batch), not on the 16M-pose step"
        if numa:
            e2e["host_affinity"] = numa
        del h_ang, h_grot, h_bone, h_root, h_gw, h_gu, out, fo

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = HBM_FALLBACK_GBS; peak_src = "fallback (B200_PROFILING.md)"

    # ---- extras, rank 0 only, never allowed to break the headline line ----
    extras = {}
    if rank == 0 and not strong:
        def guarded(key, fn):
            try:
                extras[key] = fn()
            except Exception as e:
                extras[key] = {"error": repr(e)[:300]}

        def settle():
            # the variants below are compared with the headline, so they are measured in its regime: a short burst after
            # idle, not on a board still power-throttled by whatever ran before (the clock sampler's 1 s extension, the
            # sustained loop): generator mode read 0.221 - 0.263 ms in four runs on one box without this pause
            torch.cuda.synchronize(dev)
            time.sleep(0.5)

        def generator_mode():
            half, mid = tables.generator_slot_scale(True)
            g = torch.Generator(device=dev).manual_seed(99)
            raws = [torch.randn((n, 35), generator=g, device=dev) for _ in range(2)]
            for r_ in raws:
                r_[:, 32:35] = torch.rand((n, 3), generator=g, device=dev) * 0.2 - 0.1
                r_[:, 34] += 0.1
            d_raw = torch.empty((n, 35), device=dev)
            hp, mp = half.ctypes.data, mid.ctypes.data

            def gen_step(i):
                d = sets[i % nbuf]; r_ = raws[i % 2]
                _cabi.check(lib.dhfk_generator_forward(r_.data_ptr(), 35, d["bone"].data_ptr(), 15, hp, mp, 10.0, cam_ptr,
                                                       world.data_ptr(), None, uv.data_ptr(), n, flags, sp), "gen fwd")
                _cabi.check(lib.dhfk_generator_backward(r_.data_ptr(), 35, d["bone"].data_ptr(), 15, hp, mp, 10.0, cam_ptr,
                                                        d["g_world"].data_ptr(), None, d["g_uv"].data_ptr(),
                                                        d_raw.data_ptr(), 35, n, flags, sp), "gen bwd")
            settle()
            for i in range(5):
                gen_step(i)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            gsteps = min(steps, 50)
            e0.record(stream)
            for i in range(gsteps):
                gen_step(i)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            gms = e0.elapsed_time(e1) / gsteps
            gbytes = (200 + 320) + (200 + 320 + 140)     # fwd: 50 floats in, 80 out; bwd: 50 + 80 in, 35 out
            gbs = gbytes * n / (gms * 1e-3) / 1e9
            return {"poses_per_s": n / (gms * 1e-3), "ms_per_step": gms, "bytes_per_pose": gbytes, "hbm_gbs": gbs,
                    "frac_of_copy_peak": gbs / peak,
                    "what": "dhfk_generator_forward + dhfk_generator_backward: raw network output [N,35] in, "
                            "d(raw) out; tanh / slot scatter / range map fused (SURVEY 8 f1); %d steps after 0.5 s of idle "
                            "and 5 warm-ups, like the headline" % gsteps}

        def trig_variant(ff, what):
            def run():
                settle()
                for i in range(5):
                    path1.step(i, ff)
                torch.cuda.synchronize(dev)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                fsteps = min(steps, 50)
                f0.record(stream)
                for i in range(fsteps):
                    path1.step(i, ff)
                f1.record(stream)
                torch.cuda.synchronize(dev)
                fms = f0.elapsed_time(f1) / fsteps
                return {"poses_per_s": n / (fms * 1e-3), "ms_per_step": fms,
                        "hbm_gbs": (FWD_BYTES + BWD_BYTES) * n / (fms * 1e-3) / 1e9, "what": "same step, " + what}
            return run

        def sustained():
            # the same step for >= 1.5 s: the board's power cap (sw_power_cap) pulls the SM clock down after ~0.1 s of
            # continuous work of THESE kernels (a plain copy holds its rate), so a long run sits below the burst the K-step
            # region above measures (profiles/r2e_sweep.txt).  Reported so that both regimes are on the line; the headline
            # keeps the contract (K steps after W warm-ups).
            k = max(steps, int(1.5 / max(ms_per_step * 1e-3, 1e-6)))
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(k // 4):
                path1.step(i)
            s0.record(stream)
            for i in range(k):
                path1.step(i)
            s1.record(stream)
            torch.cuda.synchronize(dev)
            sms = s0.elapsed_time(s1) / k
            # the same question for the roofline's own denominator: a plain device copy (how MEASURED_PEAKS.json's
            # hbm_gbs was taken: b.copy_(a), read + write bytes), best of 10 and held for >= 1.5 s, on this box, now
            copy = None
            try:
                a = torch.empty(1 << 29, dtype=torch.bfloat16, device=dev)      # 1 GiB each
                b = torch.empty_like(a)
                nbytes = 2 * a.numel() * a.element_size()
                best = float("inf")
                for _ in range(10):
                    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    c0.record(stream); b.copy_(a); c1.record(stream)
                    torch.cuda.synchronize(dev)
                    best = min(best, c0.elapsed_time(c1))
                reps = max(10, int(1.5e3 / best))
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(stream)
                for _ in range(reps):
                    b.copy_(a)
                c1.record(stream)
                torch.cuda.synchronize(dev)
                copy = {"burst_gbs": nbytes / (best * 1e-3) / 1e9, "sustained_gbs": nbytes * reps / (c0.elapsed_time(c1) * 1e-3) / 1e9,
                        "what": "torch b.copy_(a) over 1 GiB bf16 tensors (read + write bytes): best of 10, then %d back to back" % reps}
                del a, b
            except Exception as e:
                copy = {"error": repr(e)[:200]}
            gbs = (FWD_BYTES + BWD_BYTES) * n / (sms * 1e-3) / 1e9
            return {"poses_per_s": n / (sms * 1e-3), "ms_per_step": sms, "steps": k, "hbm_gbs": gbs,
                    "frac_of_copy_peak": gbs / peak, "device_copy_now": copy,
                    "frac_of_sustained_copy": (gbs / copy["sustained_gbs"]) if copy and "sustained_gbs" in copy else None,
                    "what": "same step, %d back-to-back steps (>= 1.5 s) after %d more as warm-up: the power-capped steady "
                            "state.  A plain device copy does NOT drop when held (device_copy_now): the gap is SM power -- "
                            "instructions per pose -- not memory" % (k, k // 4)}

        guarded("generator_mode", generator_mode)
        guarded("sustained", sustained)
        if not args.fast_trig and not args.accurate_trig:
            guarded("fast_trig_variant", trig_variant(_cabi.FLAG_FAST_TRIG, "DHFK_FLAG_FAST_TRIG: MUFU.SIN/COS in the forward too"))
            guarded("accurate_trig_variant", trig_variant(_cabi.FLAG_ACCURATE_TRIG,
                                                          "DHFK_FLAG_ACCURATE_TRIG: table sincos (abs err 1e-7) in the backward too"))
        if not args.no_extras:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            del path1, sets
            torch.cuda.empty_cache()

            def side_kernels():
                import aux_bench
                r = aux_bench.measure(dev, n, peak, steps=30)
                return {"poses": n, "what": "the kernels either side of the path (SURVEY 8 f2-f4, standalone camera ops, 32-slot "
                                            "layout), device-resident, CUDA events, 4 rotating buffer sets; frac = algorithmic "
                                            "bytes / time / measured copy peak",
                        "kernels": {k: {"ms": v["ms"], "bytes_per_pose": v["bytes_per_pose"], "frac": v["frac"]}
                                    for k, v in r["kernels"].items()}}

            def dropin_path():
                import dropin_path_bench
                return dropin_path_bench.measure(dev, peak_gbs=peak)

            guarded("side_kernels", side_kernels)
            torch.cuda.empty_cache()
            guarded("dropin_path", dropin_path)
            torch.cuda.empty_cache()
            guarded("gan_step", lambda: dict(reference_loop_timings(),
                                             what="BASELINE configs[2] (single-frame, batch 1024, dense 256) and configs[3] "
                                                  "(multi-frame, 512 clips x 9 frames, architecture 3,3): ms per iteration of "
                                                  "the reference's own loop functions (GAN_solutions_FK_generator / "
                                                  "video_mode_GAN_solutions_FK_generator, unmodified, from oracle/_ref) on this "
                                                  "GPU -- as they are, and with dhfk.dropin.install(generators, critics, "
                                                  "loader_refresh); wall clock incl. their host RNG and .cpu() copies"))

    if distributed:
        dist.barrier()
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    bwd_gbs = BWD_BYTES * n_step / (bwd_ms * 1e-3) / 1e9
    fwd_gbs = FWD_BYTES * n_step / (fwd_ms * 1e-3) / 1e9
    # The bracketed durations above include the event records around the launches (a ~2 us bubble each: fwd + bwd
    # bracketed add up to more than ms_per_step, which is dominated by the un-probed steps).  The same live shares applied
    # to the step time give durations that add up to the step: reported beside the bracketed ones, not instead of them.
    share_b = bwd_ms / (fwd_ms + bwd_ms)
    bwd_ms_share, fwd_ms_share = ms_per_step * share_b, ms_per_step * (1.0 - share_b)
    share_note = ("ms_per_launch = CUDA events around the launch in every 8th step of the timed region (includes the event "
                  "bubbles); *_step_share = ms_per_step x this kernel's share of the bracketed pair (adds up to the step)")
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")     # written from an `ncu --set full` capture (per launch)
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("dhfk_bwd_kernel_bytes_per_launch")
            traffic_src = tj.get("source", "profiles/traffic.json")
            if traffic is not None and n_step != tj.get("poses_per_launch", 1 << 20):
                traffic = traffic * n_step / tj.get("poses_per_launch", 1 << 20)
                traffic_src += " (scaled from the captured launch size to this one)"
        except Exception:
            traffic = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, {"l2": ("one input set of %.1f GB per rank, far larger than the 126 MB L2"
                                                % (n_step * (216 + 320) / 1e9)) if strong else
                                               ("inputs rotate over %d buffer sets (%.1f GB per set) > 126 MB L2"
                                                % (nbuf, n * (216 + 320) / 1e9))}),
        "roofline": {"bound": "hbm", "kernel": "dhfk_bwd_kernel<GUV=1,GBONE=0>", "achieved": bwd_gbs, "peak": peak,
                     "unit": "GB/s", "frac": bwd_gbs / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "ms_per_launch": bwd_ms, "algorithmic_bytes_per_launch": BWD_BYTES * n_step,
                     "ms_per_launch_step_share": bwd_ms_share,
                     "frac_step_share": BWD_BYTES * n_step / (bwd_ms_share * 1e-3) / 1e9 / peak, "timing": share_note},
        "roofline_fwd": {"bound": "hbm", "kernel": "dhfk_fwd_kernel<CAM=0,UV=1>", "achieved": fwd_gbs, "peak": peak,
                         "unit": "GB/s", "frac": fwd_gbs / peak, "ms_per_launch": fwd_ms,
                         "algorithmic_bytes_per_launch": FWD_BYTES * n_step, "ms_per_launch_step_share": fwd_ms_share,
                         "frac_step_share": FWD_BYTES * n_step / (fwd_ms_share * 1e-3) / 1e9 / peak},
        "roofline_step": {"achieved": (FWD_BYTES + BWD_BYTES) * n_step / (ms_per_step * 1e-3) / 1e9, "peak": peak,
                          "frac": (FWD_BYTES + BWD_BYTES) * n_step / (ms_per_step * 1e-3) / 1e9 / peak, "unit": "GB/s"},
        "clocks": dict(sampler.summary(), window="timed region" + (" + 1 s extension of the same loop" if extension else "")),
        # per rank, inside the timed region: forward + backward per step, + the exchange kernel per step when it is ours
        "gpu_launches": (3 if (strong and gbuf is not None and gbuf.peer is not None) else 2) * steps,
    }
    if parity:
        line["parity"] = parity
    if e2e:
        line["e2e"] = e2e
    if weak_extra:
        line["weak"] = weak_extra
    line.update(extras)
    if allreduce_extra:
        line["grad_allreduce"] = allreduce_extra
    if not args.no_cpu_baseline:
        res = time_reference_pipeline(args.ref_chunk, 1, steps=10, warmup=2)
        if res is not None:
            pps, dt, threads = res
            line["cpu_baseline"] = {"value": pps, "unit": UNIT, "cores": threads, "kind": "reference",
                                    "sample": "10 x %d-pose calls (reference batch size), fwd+bwd, the unmodified reference "
                                              "from oracle/_ref/dh_aug_ref.zip" % args.ref_chunk}
        pps, dt, threads = time_torch_port(args.ref_chunk, 1, steps=10, warmup=2)
        port = {"value": pps, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "10 x %d-pose torch calls (reference batch size), fwd+bwd, oracle/torch_port.py" % args.ref_chunk}
        line["cpu_baseline_port" if "cpu_baseline" in line else "cpu_baseline"] = port
        try:
            fps, fthreads = time_torch_port_forward_only(args.ref_chunk)
            line["cpu_baseline_forward_only"] = {"value": fps, "unit": UNIT, "cores": fthreads, "kind": "port",
                                                 "sample": "best of 5 %d-pose torch calls under no_grad (forward only)" % args.ref_chunk}
        except Exception as e:
            line["cpu_baseline_forward_only"] = {"error": repr(e)}
        try:
            gps, gdt = time_torch_port_gpu(dev)
            line["torch_eager_gpu"] = {"value": gps, "unit": UNIT, "ms_per_call": gdt * 1e3, "kind": "port",
                                       "sample": "16384-pose torch-eager calls on this GPU, fwd+bwd, best of 3 "
                                                 "(the reference's op sequence, oracle/torch_port.py; context, not an arm)"}
        except Exception as e:
            line["torch_eager_gpu"] = {"error": repr(e)}
        try:
            cps, cthreads = time_c_oracle()
            line["cpu_baseline_c"] = {"value": cps, "unit": UNIT, "cores": cthreads, "kind": "port",
                                      "sample": "131072 poses fwd+bwd, oracle/dhfk_oracle.c (float64, OpenMP)"}
        except Exception as e:
            line["cpu_baseline_c"] = {"error": repr(e)}
    emit_json(line)
    if distributed:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
