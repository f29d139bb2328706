"""Data-parallel plumbing for the augmentation path.

Poses (rows) are independent, so rank r of R simply takes rows [r*N/R, (r+1)*N/R) -- there is no
collective on the FK / projection data path.  The only exchange in a GAN step is the gradient
all-reduce of the small generator / critic MLPs; ``allreduce_grads_flat`` sends each model's
gradients as ONE flat buffer (1-4 MB => latency-bound over NVLink/NVSwitch; one launch instead of
one per parameter).  Works with the nccl backend on GPUs and with gloo on CPU (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(n: int, rank: int, world_size: int):
    """Contiguous row range [lo, hi) of rank `rank`; sizes differ by at most one row."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %d/%d" % (rank, world_size))
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_grads_flat(params, group=None, average: bool = True):
    """Sum (or average) the .grad of `params` across ranks with a single all-reduce."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    views, off = [], 0
    for g in grads:
        k = g.numel()
        views.append(flat[off:off + k].view_as(g))
        off += k
    torch._foreach_copy_(grads, views)      # one multi-tensor launch instead of one copy per parameter
    return flat.numel()


class FlatGradBuffer:
    """ONE persistent flat fp32 buffer whose slices ARE the ``.grad`` tensors of the given parameters.

    The data-parallel GAN step has five optimizer steps per iteration (3-D critic, its flip pass, 2-D critic, its flip
    pass, generator: model_fk_gan_train.py:314-341, 382-409, 415-482), each of which needs the ranks' gradients
    averaged.  With the gradients living inside one buffer the exchange is a single NCCL all-reduce on that buffer:
    no ``torch.cat`` to pack, no copy back, no per-parameter launches (``allreduce_grads_flat`` above costs
    cat + all-reduce + divide + foreach-copy: 0.14 ms for 2 MB on 8 B200s, most of it launches).  Several models can
    share one buffer (``FlatGradBuffer([*G.parameters(), *D3.parameters(), *D2.parameters()])``) and be reduced in one
    call, or separately through ``allreduce(span=buf.span_of(model))``.

    ``peer_exchange=True`` (NCCL process group, one node): the buffer is a symmetric allocation and ``allreduce`` over
    the group it was created for is ``dhfk_grad_allreduce`` -- one hand-written kernel over NVLink peer memory / the
    NVSwitch multicast object, in place (see ``PeerExchange``); where that is not available the buffer is an ordinary
    tensor and the same call is NCCL's.

    Zero the gradients with ``buf.zero()`` (or ``zero_grad(set_to_none=False)``): ``zero_grad()``'s default drops the
    ``.grad`` tensors, after which autograd allocates fresh ones outside the buffer.  ``allreduce`` notices and re-adopts
    such gradients (one copy each), so the result is always right; it is only fastest when nothing was dropped.
    """

    def __init__(self, params, device=None, dtype=torch.float32, peer_exchange=False, group=None, max_ctas=16,
                 cta_threads=512, timeout_ms=20000):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters that require grad")
        device = device or self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4            # every slice starts 16-byte aligned
        self.peer, self._peer_group = None, group
        if peer_exchange and dtype == torch.float32:
            self.peer = PeerExchange.create(off, device, group, max_ctas=max_ctas, cta_threads=cta_threads,
                                            timeout_ms=timeout_ms)
        if self.peer is not None:
            self.flat = self.peer.buffer               # symmetric allocation: every rank's buffer is mapped in every rank
        else:
            self.flat = torch.zeros(off, dtype=dtype, device=device)
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(self.offsets, self.params)]
        self.attach()

    def attach(self):
        """Point every parameter's .grad at its slice (keeps values already accumulated elsewhere)."""
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v
        return self

    def zero(self):
        self.flat.zero_()
        for p, v in zip(self.params, self.views):     # re-adopt anything zero_grad(set_to_none=True) dropped
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def span_of(self, module_or_params):
        """(lo, hi) element range of the buffer covering the given module's parameters (must be contiguous in it)."""
        want = {id(p) for p in (module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params)}
        idx = [i for i, p in enumerate(self.params) if id(p) in want]
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            raise ValueError("parameters are not a contiguous run of this buffer")
        return self.offsets[idx[0]], self.offsets[idx[-1]] + (self.params[idx[-1]].numel() + 3) // 4 * 4

    def allreduce(self, group=None, average=True, span=None):
        """Average (or sum) the gradients across ranks with one collective on the buffer (or on `span` of it).
        Returns the number of elements sent.  Runs on the current stream's semantics like any torch collective: call
        it under ``torch.cuda.stream(side)`` to overlap it with kernels on the main stream."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return 0
        for p, v in zip(self.params, self.views):     # gradients that escaped the buffer: adopt them (correct, slower)
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        if self.peer is not None and group is self._peer_group:      # the ranks the symmetric buffer was set up over
            lo, hi = (0, self.flat.numel()) if span is None else span
            return self.peer.allreduce(lo, hi, average)
        buf = self.flat if span is None else self.flat[span[0]:span[1]]
        if average and "nccl" in str(dist.get_backend(group)):
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group)      # the division happens inside the collective
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            if average:
                buf /= dist.get_world_size(group)
        return buf.numel()


class PeerExchange:
    """The gradient exchange as ONE hand-written kernel over NVLink peer memory (csrc/dhfk_allreduce.cu,
    ``dhfk_grad_allreduce``): every rank reduces its slice of the buffer -- through the NVSwitch's multicast object
    (NVLS: the switch adds the copies in flight and replicates the result) when the node has one, otherwise with loads
    from / stores to every peer -- in place, with two cross-GPU flag barriers inside the kernel.  PyTorch only provides
    the memory: ``torch.distributed._symmetric_memory`` allocates the buffer and the flag block and maps every rank's
    copy into every process.  ``create`` returns None where that is not possible (one rank, gloo, no peer access), and
    the caller keeps NCCL.  ``timeout_ms`` bounds how long a rank's kernel waits for the slowest rank to reach the same
    call (20 s by default: ranks of a training loop drift apart by far less; a rank that never arrives must not hang
    the GPU) -- ``check()`` tells whether that ever happened."""

    def __init__(self, buffer, flags, status, bufs, flag_ptrs, mc_ptr, rank, world, max_ctas, cta_threads, timeout_ms,
                 handles):
        import ctypes
        self.buffer, self.flags, self.status = buffer, flags, status
        self.rank, self.world, self.timeout_ms = rank, world, int(timeout_ms)
        self.max_ctas, self.cta_threads = int(max_ctas), int(cta_threads)
        self.multicast = bool(mc_ptr)
        self._mc_ptr = int(mc_ptr or 0)
        self._buf_ptrs = [int(b) for b in bufs]
        self._flag_arr = (ctypes.c_void_p * world)(*[int(f) for f in flag_ptrs])
        self._ranges = {}                # lo -> ctypes array of every rank's address of element lo
        self._handles = handles          # keeps the mappings alive
        self._status_ptr = status.data_ptr()
        self._dev_index = buffer.device.index
        self.use_multicast = self.multicast
        self.peer_min_ctas = 32

    @classmethod
    def create(cls, numel, device, group=None, max_ctas=16, cta_threads=512, timeout_ms=20000):
        from . import _cabi
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        if "nccl" not in str(dist.get_backend(group)) or dist.get_world_size(group) > _cabi.AR_MAX_WORLD:
            return None
        px = None
        try:
            import torch.distributed._symmetric_memory as symm
            pg = group if group is not None else dist.group.WORLD
            numel = (int(numel) + 3) // 4 * 4
            buf = symm.empty(numel, dtype=torch.float32, device=device)
            buf.zero_()
            flags = symm.empty(_cabi.AR_FLAG_WORDS, dtype=torch.int32, device=device)
            flags.zero_()
            torch.cuda.synchronize(device)
            hb = symm.rendezvous(buf, pg)
            hf = symm.rendezvous(flags, pg)
            hf.barrier()                              # every rank's flags are zero before anyone signals
            status = torch.zeros(1, dtype=torch.int32, device=device)
            mc = int(getattr(hb, "multicast_ptr", 0) or 0)
            px = cls(buf, flags, status, list(hb.buffer_ptrs), list(hf.buffer_ptrs), mc, hb.rank, hb.world_size,
                     max_ctas, cta_threads, timeout_ms, (hb, hf))
        except Exception as e:           # no peer access / no symmetric-memory support: the caller keeps NCCL
            cls.last_error = repr(e)
        # all ranks or none: a rank on the kernel and a rank on NCCL would wait for each other forever.  The multicast
        # mapping must agree too (a rank without it would take the peer path while the others use the switch).
        try:
            state = torch.tensor([1 if px is not None else 0, 1 if (px is not None and px.multicast) else 0],
                                 dtype=torch.int32, device=device)
            dist.all_reduce(state, op=dist.ReduceOp.MIN, group=group)
            everyone, everyone_mc = (int(v) for v in state.tolist())
        except Exception as e:
            cls.last_error = repr(e)
            return None
        if not everyone:
            if px is not None:
                cls.last_error = "symmetric memory unavailable on another rank"
            return None
        if not everyone_mc:
            px.multicast = px.use_multicast = False
        return px

    last_error = None

    def allreduce(self, lo=0, hi=None, average=True):
        """All-reduce elements [lo, hi) of the buffer (multiples of 4) on the current stream.  Returns hi - lo."""
        from . import _cabi
        hi = self.buffer.numel() if hi is None else hi
        if lo % 4 or hi % 4 or not (0 <= lo <= hi <= self.buffer.numel()):
            raise ValueError("range must be 16-byte aligned and inside the buffer")
        if hi == lo:
            return 0
        arr = self._ranges.get(lo)
        if arr is None:
            import ctypes
            arr = self._ranges[lo] = (ctypes.c_void_p * self.world)(*[b + 4 * lo for b in self._buf_ptrs])
        mc = self._mc_ptr + 4 * lo if (self.use_multicast and self._mc_ptr) else None
        # without the switch every element is `world` round trips: the peer path wants twice the CTAs (6.4 MB on 2 / 8
        # GPUs: 66 us with 16 CTAs, 41 / 49 us with 32; the NVLS path is flat from 8 CTAs on)
        ctas = self.max_ctas if mc is not None else min(max(self.max_ctas, self.peer_min_ctas), _cabi.AR_MAX_CTAS)
        rc = _cabi.load().dhfk_grad_allreduce(arr, mc, self._flag_arr, self._status_ptr, self.rank, self.world, hi - lo,
                                              (1.0 / self.world) if average else 1.0, ctas,
                                              self.cta_threads, self.timeout_ms,
                                              torch._C._cuda_getCurrentRawStream(self._dev_index))
        if rc:
            _cabi.check(rc, "dhfk_grad_allreduce")
        return hi - lo

    def check(self):
        """Raise if any exchange so far gave up waiting for a peer (synchronises)."""
        st = int(self.status.item())
        if st:
            raise RuntimeError("dhfk_grad_allreduce: exchange %d timed out waiting for a peer rank" % st)


def grad_allreduce_group(max_ctas=4):
    """A dedicated NCCL communicator for the gradient exchange, limited to `max_ctas` CTAs.

    The gradients of the GAN's three MLPs are a few MB: latency-bound, a handful of CTAs move them at NVLink speed.
    NCCL's default launch takes 16-32 CTAs -- a fifth of the SMs -- and holds them while it waits for its peers, which is
    what kernels running beside it on another stream pay for.  Measured on 4 B200s (profiles/r2k_nccl_sweep.txt), 6.4 MB:
    default communicator 45 us alone; max_ctas 8 / 4 / 2: 94 / 153 / 283 us.  Behind a 0.85 ms FK step all of them hide
    completely and the FK kernels run 3 % faster beside the small ones; behind a 0.2 ms step only the default one fits.
    So this is an option for long overlapped steps, not the default (bench.py --nccl-max-ctas).  Returns None (= the
    default group) off NCCL or when the installed torch cannot configure it."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_backend() != "nccl" or not max_ctas:
        return None
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(max_ctas)
        opts.config.min_ctas = 1
        return dist.new_group(backend="nccl", pg_options=opts)
    except Exception:
        return None


def broadcast_camera_choice(subject_id: int, cam_id: int, device, src: int = 0, group=None):
    """All ranks must project with the same (subject, camera) per iteration, as the reference does per
    batch (model_fk_gan_train.py:344-347): rank `src` draws, everyone receives."""
    t = torch.tensor([subject_id, cam_id], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=src, group=group)
    return int(t[0].item()), int(t[1].item())
