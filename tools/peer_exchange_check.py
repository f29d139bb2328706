#!/usr/bin/env python
"""dhfk_grad_allreduce (csrc/dhfk_allreduce.cu) against NCCL, under torchrun on >= 2 GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tools/peer_exchange_check.py [--time]

Checks (every rank; rank 0 prints): the average / sum equals NCCL's within fp32 rounding of the sum order, all ranks
end bit-identical, sub-ranges leave the rest of the buffer untouched, repeated calls (flag counters) keep working, both
the NVLS multicast path and the plain peer path.  --time adds the 6.4 MB timing (alone, back to back) beside NCCL.
Exit code 0 = all passed, 3 = symmetric memory not available on this box (nothing to check)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dhfk import parallel
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    numel = 1590236 // 4 * 4 + 8          # the bench's generator + two critics, plus a tail
    px = parallel.PeerExchange.create(numel, dev, max_ctas=16, timeout_ms=2000)
    ok_all = torch.tensor([1 if px is not None else 0], device=dev)
    dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
    if not int(ok_all.item()):
        if rank == 0:
            print(json.dumps({"peer_exchange": "unavailable", "why": parallel.PeerExchange.last_error}))
        dist.destroy_process_group()
        return 3
    out = {"world": world, "multicast": px.multicast, "numel": numel, "cases": []}
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    failures = []

    def one_case(lo, hi, average, use_mc, tag):
        px.use_multicast = use_mc
        data = torch.randn(numel, generator=g, device=dev) * (1.0 + rank)
        px.buffer.copy_(data)
        want = data.clone()
        part = want[lo:hi].clone()
        dist.all_reduce(part, op=dist.ReduceOp.AVG if average else dist.ReduceOp.SUM)
        want[lo:hi] = part
        torch.cuda.synchronize(dev)
        dist.barrier()
        px.allreduce(lo, hi, average)
        torch.cuda.synchronize(dev)
        px.check()
        got = px.buffer.clone()
        err = float((got - want).abs().max())
        scale = float(want.abs().max())
        # identical on every rank?
        ref = got[lo:hi].clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, got[lo:hi]))
        untouched = bool(torch.equal(got[:lo], data[:lo]) and torch.equal(got[hi:], data[hi:]))
        rec = {"case": tag, "lo": lo, "hi": hi, "average": average, "multicast": bool(use_mc and px.multicast),
               "max_abs_err_vs_nccl": err, "scale": scale, "bit_identical_across_ranks": same, "rest_untouched": untouched}
        out["cases"].append(rec)
        if not (err <= 2e-6 * max(scale, 1.0) * world and same and untouched):
            failures.append(rec)

    for use_mc in ((True, False) if px.multicast else (False,)):
        one_case(0, numel, True, use_mc, "whole buffer, average")
        one_case(0, numel, False, use_mc, "whole buffer, sum")
        one_case(4096, 4096 + 440000 // 4 * 4, True, use_mc, "one model's span")
        one_case(numel - 8, numel, True, use_mc, "8-element tail")
        one_case(16, 20, True, use_mc, "one 16-byte element")
        for _ in range(20):               # flag counters over many calls, no host sync in between
            px.allreduce(0, numel, True)
        torch.cuda.synchronize(dev)
        px.check()
        one_case(0, numel, True, use_mc, "after 20 back-to-back calls")

    # the exchange captured in a CUDA graph and replayed: after the first replay the ranks are identical, and averaging
    # identical values is exact, so five replays still equal NCCL's average of the initial data
    px.use_multicast = px.multicast
    data = torch.randn(numel, generator=g, device=dev) * (1.0 + rank)
    want = data.clone()
    dist.all_reduce(want, op=dist.ReduceOp.AVG)
    px.buffer.copy_(data)
    torch.cuda.synchronize(dev)
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        px.allreduce(0, numel, True)
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize(dev)
    px.check()
    err = float((px.buffer - want).abs().max())
    rec = {"case": "captured in a CUDA graph, replayed 5x", "max_abs_err_vs_nccl": err, "multicast": px.multicast}
    out["cases"].append(rec)
    if not err <= 2e-6 * float(want.abs().max()) * world:
        failures.append(rec)
    del graph

    # end to end: a data-parallel gradient through FlatGradBuffer(peer_exchange=True) equals the full-batch gradient
    # (the same check tests/test_parallel_gloo.py makes on CPU), including re-adoption after zero_grad() and a span
    torch.manual_seed(0)
    G = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 5)).to(dev)
    D = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 1)).to(dev)
    nrow = 64 * world
    x = torch.randn(nrow, 8, generator=torch.Generator().manual_seed(3)).to(dev)
    lo, hi = parallel.shard_rows(nrow, rank, world)
    fb = parallel.FlatGradBuffer(list(G.parameters()) + list(D.parameters()), peer_exchange=True)
    dp_ok = fb.peer is not None
    worst = 0.0
    for it in range(3):
        if it == 1:
            G.zero_grad(); D.zero_grad()              # drops the .grad tensors: the buffer must re-adopt them
        else:
            fb.zero()
        (D(G(x[lo:hi])).sum() / nrow * world).backward()
        fb.allreduce(average=True)
        refG = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 5)).to(dev)
        refD = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 1)).to(dev)
        refG.load_state_dict(G.state_dict()); refD.load_state_dict(D.state_dict())
        (refD(refG(x)).sum() / nrow).backward()
        for a, b in zip(list(G.parameters()) + list(D.parameters()), list(refG.parameters()) + list(refD.parameters())):
            worst = max(worst, float((a.grad - b.grad).abs().max()))
            dp_ok = dp_ok and a.grad.data_ptr() == fb.views[[id(q) for q in fb.params].index(id(a))].data_ptr()
    fb.zero()
    (D(G(x[lo:hi])).sum() * (rank + 1)).backward()
    before_G = [q.grad.clone() for q in G.parameters()]
    fb.allreduce(average=False, span=fb.span_of(D))    # one model's span: the other model's gradients stay local
    torch.cuda.synchronize(dev)
    fb.peer.check()
    dp_ok = dp_ok and all(torch.equal(a, q.grad) for a, q in zip(before_G, G.parameters()))
    rec = {"case": "FlatGradBuffer(peer_exchange=True): data-parallel gradient vs full batch", "max_abs_err": worst,
           "aliased_and_span_ok": bool(dp_ok)}
    out["cases"].append(rec)
    if not (worst < 1e-6 and dp_ok):
        failures.append(rec)

    if "--time" in sys.argv:
        timing = {}
        buf_nccl = torch.randn(numel, device=dev)

        import time

        def timed(fn, reps=200):
            for _ in range(10):
                fn()
            torch.cuda.synchronize(dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            host = (time.perf_counter() - t0) / reps * 1e6      # time to ENQUEUE one call: GPU-bound only if well below gpu
            torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3, host], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return {"gpu_us": round(float(t[0]), 2), "host_enqueue_us": round(float(t[1]), 2)}

        px.peer_min_ctas = 1          # the sweep below sets the CTA count itself
        timing["nccl"] = timed(lambda: dist.all_reduce(buf_nccl, op=dist.ReduceOp.AVG))
        for use_mc in ((True, False) if px.multicast else (False,)):
            px.use_multicast = use_mc
            for threads, ctas in ((512, 8), (512, 16), (512, 32), (256, 32), (256, 64), (128, 32), (128, 64)):
                px.max_ctas, px.cta_threads = ctas, threads
                px.buffer.normal_()
                timing["%s_%dx%d" % ("multimem" if use_mc else "peer", ctas, threads)] = timed(lambda: px.allreduce(0, numel, True))
        px.use_multicast, px.max_ctas, px.cta_threads = px.multicast, 16, 512
        px.check()
        out["timing_6p4MB"] = timing
        # latency vs bandwidth: the same call over growing ranges (16 B = the two barriers and the launch, nothing else)
        sizes = {}
        for use_mc in ((True, False) if px.multicast else (False,)):
            px.use_multicast = use_mc
            px.max_ctas, px.cta_threads = (16, 512) if use_mc else (32, 512)
            for nfl in (4, 16384, 262144, 1048576, numel):
                nfl = min(nfl, numel) // 4 * 4
                sizes["%s_%d_bytes" % ("multimem" if use_mc else "peer", nfl * 4)] = timed(lambda: px.allreduce(0, nfl, True))
        for nfl in (4, 16384, 262144, 1048576, numel):
            nfl = min(nfl, numel) // 4 * 4
            sizes["nccl_%d_bytes" % (nfl * 4)] = timed(lambda: dist.all_reduce(buf_nccl[:nfl], op=dist.ReduceOp.AVG))
        px.use_multicast, px.max_ctas, px.cta_threads = px.multicast, 16, 512
        px.check()
        out["timing_by_size"] = sizes

    out["failures"] = failures
    if rank == 0:
        print(json.dumps(out, indent=1))
    bad = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX)
    dist.barrier()
    dist.destroy_process_group()
    return 1 if int(bad.item()) else 0


if __name__ == "__main__":
    sys.exit(main())
