#!/usr/bin/env python
"""bench.py -- augmented poses/sec of the fused DH-FK + projection forward+backward path.

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch of synthetic poses: dhfk_forward
(world16 + uv16) followed by dhfk_backward (d angles, d global rotation, d root from upstream
gradients on world16 and uv16).  Workload = BASELINE.json configs[1]: 1,048,576 poses per GPU,
S1/cam0, generator-range angles, template bone lengths x (1 +- 0.2), in-volume roots.
Rows shard across ranks with no data-path collective (weak scaling: per-GPU batch fixed).

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
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {
        "workload": "BASELINE configs[1]: fused DH-FK + projection forward+backward, 1M poses per B200, "
                    "16-joint H36M skeleton, synthetic generator-range angles + template bone lengths, S1/cam0",
        "poses_per_gpu": args.poses,
        "outputs": "world16[N,16,3]+uv16[N,16,2]; grads d_ang[N,33]+d_grot[N,3]+d_root[N,3]",
        "bytes_per_pose": {"forward": FWD_BYTES, "backward": BWD_BYTES},
        "parallelism": "dp%d (rows sharded, no data-path collective)" % args.gpus,
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    # bound the run: a step is ref_chunks x ref_chunk poses (~0.1 s per 1024-pose call on 8 cores)
    pps, dt, threads = time_torch_port(args.ref_chunk, args.ref_chunks, steps, warmup)
    sample = "%d x %d-pose torch calls per step (reference batch size), fwd+bwd, CPU" % (args.ref_chunks, args.ref_chunk)
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, {"reference_arm": "oracle/torch_port.py: the reference's own torch op sequence "
                                                           "(bit-identical to /root/reference on the build host), host cores only"}),
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# ---------------------------------------------------------------------------------------------------
def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dhfk
    from dhfk import _cabi, synthetic, tables

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

    lib = _cabi.load()
    n = args.poses
    flags = (_cabi.FLAG_FAST_TRIG if args.fast_trig else 0) | (_cabi.FLAG_ACCURATE_TRIG if args.accurate_trig else 0)
    blk = tables.camera_block("S1", 0)
    cam_ptr = blk.ctypes.data
    nbuf = max(1, args.buffers)
    sets = []
    for b in range(nbuf):
        d = synthetic.gan_like_torch(n, dev, seed=1234 + 97 * rank + b)
        g = torch.Generator(device=dev).manual_seed(4321 + 97 * rank + b)
        d["g_world"] = torch.randn((n, 16, 3), generator=g, device=dev)
        d["g_uv"] = torch.randn((n, 16, 2), generator=g, device=dev)
        sets.append(d)
    world = torch.empty((n, 16, 3), device=dev); uv = torch.empty((n, 16, 2), device=dev)
    g_ang = torch.empty((n, 33), device=dev); g_grot = torch.empty((n, 3), device=dev)
    g_root = torch.empty((n, 3), device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    def fwd(d):
        rc = lib.dhfk_forward(d["ang"].data_ptr(), 33, d["grot"].data_ptr(), 3, d["bone"].data_ptr(), 15,
                              d["root"].data_ptr(), 3, cam_ptr, None, 0, world.data_ptr(), None, uv.data_ptr(),
                              n, flags, sp)
        _cabi.check(rc, "dhfk_forward")

    def bwd(d):
        rc = lib.dhfk_backward(d["ang"].data_ptr(), 33, d["grot"].data_ptr(), 3, d["bone"].data_ptr(), 15,
                               d["root"].data_ptr(), 3, cam_ptr, None, 0, d["g_world"].data_ptr(), None,
                               d["g_uv"].data_ptr(), g_ang.data_ptr(), 33, g_grot.data_ptr(), 3, g_root.data_ptr(), 3,
                               None, 15, n, flags, sp)
        _cabi.check(rc, "dhfk_backward")

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # parity before any timing is reported: the first rows of buffer set 0 against the float64 C oracle,
    # as the checker only (part of the cpu_baseline leg: skipped with --no-cpu-baseline)
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        import c_oracle
        m = min(n, 4096)
        d = sets[0]
        fwd(d); bwd(d)
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
    for i in range(warmup):
        fwd(sets[i % nbuf]); bwd(sets[i % nbuf])
    barrier()

    sampler = ClockSampler(physical_gpu_index(local_rank)) if rank == 0 else None
    if sampler:
        sampler.start()
    # One event pair brackets the K timed steps; every `probe_every`-th step additionally carries three events around its
    # two launches for the live per-kernel durations (an event record is a ~2 us bubble on the stream, so probing every
    # launch would cost the headline ~3 %).
    probe_every = 8 if steps >= 16 else 1
    probes = {i: [torch.cuda.Event(enable_timing=True) for _ in range(3)] for i in range(0, steps, probe_every)}
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev_start.record(stream)
    for i in range(steps):
        d = sets[i % nbuf]
        pr = probes.get(i)
        if pr is None:
            fwd(d); bwd(d)
        else:
            pr[0].record(stream); fwd(d)
            pr[1].record(stream); bwd(d)
            pr[2].record(stream)
    ev_end.record(stream)
    torch.cuda.synchronize(dev)
    total_ms = ev_start.elapsed_time(ev_end)
    barrier()
    # keep the sampler fed if the timed region was shorter than a few polling periods
    extension = False
    if sampler and len(sampler.samples) < 10:
        extension = True
        t_end = time.time() + 1.0
        i = 0
        while time.time() < t_end:
            fwd(sets[i % nbuf]); bwd(sets[i % nbuf]); i += 1
            if i % 32 == 0:
                torch.cuda.synchronize(dev)
        torch.cuda.synchronize(dev)
    if sampler:
        sampler.stop()
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in probes.values()) / len(probes)
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in probes.values()) / len(probes)

    t = torch.tensor([total_ms, fwd_ms, bwd_ms], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, fwd_ms, bwd_ms = (float(x) for x in t.tolist())
    ms_per_step = total_ms / steps
    value = n * world_size / (ms_per_step * 1e-3)

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
        chunk, slots = 1 << 17, 3
        for _ in range(2):
            out = dhfk.fk_project_host(h_ang, h_grot, h_bone, h_root, blk, h_gw, h_gu, chunk_rows=chunk, num_streams=slots,
                                       workspace=out.get("_workspace"), out=out, fast_trig=args.fast_trig,
                                       accurate_grad=args.accurate_trig)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = dhfk.fk_project_host(h_ang, h_grot, h_bone, h_root, blk, h_gw, h_gu, chunk_rows=chunk, num_streams=slots,
                                       workspace=out["_workspace"], out=out, fast_trig=args.fast_trig,
                                       accurate_grad=args.accurate_trig)
        torch.cuda.synchronize(dev)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if distributed:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        e2e = {"value": n * world_size / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * (FWD_BYTES - 320 + 320),
               "d2h_bytes_per_step": n * (320 + 156), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "api": "dhfk.fk_project_host -> dhfk_forward_backward_host (pinned host buffers, %d-row chunks through an "
                      "upload / compute / download stream pipeline over %d device slots)" % (chunk, slots)}
        if numa:
            e2e["host_affinity"] = numa

    # ---- extra: generator-epilogue mode, same batch, device-resident, rank 0 only ----
    gen_extra = None
    if rank == 0:
        try:
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
            gen_extra = {"poses_per_s": n / (gms * 1e-3), "ms_per_step": gms, "bytes_per_pose": gbytes,
                         "hbm_gbs": gbytes * n / (gms * 1e-3) / 1e9,
                         "what": "dhfk_generator_forward + dhfk_generator_backward: raw network output [N,35] in, "
                                 "d(raw) out; tanh / slot scatter / range map fused (SURVEY 8 f1)"}
        except Exception as e:      # never let the extra break the headline line
            gen_extra = {"error": repr(e)}

    # ---- extra: the same step under the two other trig policies of include/dhfk.h (rank 0 only, not the headline;
    # distances to exact arithmetic and to the reference: tools/trig_parity.py) ----
    trig_extra = None
    if rank == 0 and not args.fast_trig and not args.accurate_trig:
        trig_extra = {}
        for key, ff, what in (("fast_trig_variant", _cabi.FLAG_FAST_TRIG, "DHFK_FLAG_FAST_TRIG: MUFU.SIN/COS in the forward too"),
                              ("accurate_trig_variant", _cabi.FLAG_ACCURATE_TRIG,
                               "DHFK_FLAG_ACCURATE_TRIG: table sincos (abs err 1e-7) in the backward too")):
            try:
                def var_step(i, ff=ff):
                    d = sets[i % nbuf]
                    _cabi.check(lib.dhfk_forward(d["ang"].data_ptr(), 33, d["grot"].data_ptr(), 3, d["bone"].data_ptr(), 15,
                                                 d["root"].data_ptr(), 3, cam_ptr, None, 0, world.data_ptr(), None,
                                                 uv.data_ptr(), n, ff, sp), "fwd")
                    _cabi.check(lib.dhfk_backward(d["ang"].data_ptr(), 33, d["grot"].data_ptr(), 3, d["bone"].data_ptr(), 15,
                                                  d["root"].data_ptr(), 3, cam_ptr, None, 0, d["g_world"].data_ptr(), None,
                                                  d["g_uv"].data_ptr(), g_ang.data_ptr(), 33, g_grot.data_ptr(), 3,
                                                  g_root.data_ptr(), 3, None, 15, n, ff, sp), "bwd")
                for i in range(5):
                    var_step(i)
                torch.cuda.synchronize(dev)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                fsteps = min(steps, 50)
                f0.record(stream)
                for i in range(fsteps):
                    var_step(i)
                f1.record(stream)
                torch.cuda.synchronize(dev)
                fms = f0.elapsed_time(f1) / fsteps
                trig_extra[key] = {"poses_per_s": n / (fms * 1e-3), "ms_per_step": fms,
                                   "hbm_gbs": (FWD_BYTES + BWD_BYTES) * n / (fms * 1e-3) / 1e9, "what": "same step, " + what}
            except Exception as e:
                trig_extra[key] = {"error": repr(e)}

    # ---- extra (N > 1): the only exchange of the data-parallel GAN step, the flat gradient all-reduce of a
    # generator/critic-sized model (1-4 MB, latency-bound), outside the timed region ----
    allreduce_extra = None
    if distributed:
        from dhfk import parallel
        net = torch.nn.Sequential(torch.nn.Linear(128, 256), *[torch.nn.Linear(256, 256) for _ in range(7)],
                                  torch.nn.Linear(256, 35)).to(dev)           # Fk_Generator-sized (dense 256)
        for p_ in net.parameters():
            p_.grad = torch.ones_like(p_)
        for _ in range(3):
            parallel.allreduce_grads_flat(list(net.parameters()))
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(20):
            nel = parallel.allreduce_grads_flat(list(net.parameters()))
        a1.record(stream)
        torch.cuda.synchronize(dev)
        ta = torch.tensor([a0.elapsed_time(a1) / 20], device=dev, dtype=torch.float64)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        allreduce_extra = {"bytes": int(nel) * 4, "ms": float(ta.item()), "backend": "nccl",
                           "what": "dhfk.parallel.allreduce_grads_flat of a dense-256 generator's gradients "
                                   "(cat + one all-reduce + scatter back); not part of the timed FK region"}

    if distributed:
        dist.barrier()
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = HBM_FALLBACK_GBS; peak_src = "fallback (B200_PROFILING.md)"
    bwd_gbs = BWD_BYTES * n / (bwd_ms * 1e-3) / 1e9
    fwd_gbs = FWD_BYTES * n / (fwd_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")     # written from an `ncu --set full` capture (per launch)
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dhfk_bwd_kernel_bytes_per_launch")
        except Exception:
            traffic = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, {"l2": "inputs rotate over %d buffer sets (%.1f GB per set) > 126 MB L2"
                                                % (nbuf, n * (216 + 320) / 1e9)}),
        "roofline": {"bound": "hbm", "kernel": "dhfk_bwd_kernel<GUV=1,GBONE=0>", "achieved": bwd_gbs, "peak": peak,
                     "unit": "GB/s", "frac": bwd_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                     "ms_per_launch": bwd_ms, "algorithmic_bytes_per_launch": BWD_BYTES * n},
        "roofline_fwd": {"bound": "hbm", "kernel": "dhfk_fwd_kernel<CAM=0,UV=1>", "achieved": fwd_gbs, "peak": peak,
                         "unit": "GB/s", "frac": fwd_gbs / peak, "ms_per_launch": fwd_ms,
                         "algorithmic_bytes_per_launch": FWD_BYTES * n},
        "roofline_step": {"achieved": (FWD_BYTES + BWD_BYTES) * n / (ms_per_step * 1e-3) / 1e9, "peak": peak,
                          "frac": (FWD_BYTES + BWD_BYTES) * n / (ms_per_step * 1e-3) / 1e9 / peak, "unit": "GB/s"},
        "clocks": dict(sampler.summary(), window="timed region" + (" + 1 s extension of the same loop" if extension else "")),
        "gpu_launches": 2 * steps,
    }
    if parity:
        line["parity"] = parity
    if e2e:
        line["e2e"] = e2e
    if gen_extra:
        line["generator_mode"] = gen_extra
    if trig_extra:
        line.update(trig_extra)
    if allreduce_extra:
        line["grad_allreduce"] = allreduce_extra
    if not args.no_cpu_baseline:
        pps, dt, threads = time_torch_port(args.ref_chunk, 1, steps=10, warmup=2)
        line["cpu_baseline"] = {"value": pps, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "10 x %d-pose torch calls (reference batch size), fwd+bwd, oracle/torch_port.py" % args.ref_chunk}
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
