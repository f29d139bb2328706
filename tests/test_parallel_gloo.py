"""N>1 host path on CPU: world_size-2 gloo.  Rows are sharded with no data-path collective; the only
exchange is the flat gradient all-reduce of the (tiny) GAN models and the camera-choice broadcast."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dhfk import parallel
        n = 1001
        lo, hi = parallel.shard_rows(n, rank, world)
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        x = torch.randn(n, 8)
        ((model(x[lo:hi]) ** 2).sum() / n).backward()          # each rank: its shard of the global mean
        sent = parallel.allreduce_grads_flat(model.parameters(), average=False)
        ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        ref.load_state_dict(model.state_dict())
        ((ref(x) ** 2).sum() / n).backward()
        err = max((a.grad - b.grad).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
        subj, cam = parallel.broadcast_camera_choice(3 if rank == 0 else 0, 2 if rank == 0 else 1, "cpu")
        counts = torch.tensor([hi - lo])
        dist.all_reduce(counts)
        q.put((rank, err, sent, subj, cam, int(counts.item())))
    finally:
        dist.destroy_process_group()


def test_shard_and_flat_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    nparams = 8 * 16 + 16 + 16 * 4 + 4
    for rank, err, sent, subj, cam, total in res:
        assert err < 1e-6
        assert sent == nparams          # ONE flat buffer carrying every gradient
        assert (subj, cam) == (3, 2)    # everyone projects with rank 0's camera
        assert total == 1001


def _flatbuf_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dhfk import parallel
        n = 640
        lo, hi = parallel.shard_rows(n, rank, world)
        torch.manual_seed(0)
        G = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 5))
        D = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 1))
        x = torch.randn(n, 8)
        # peer_exchange asks for the NVLink kernel; off NCCL / GPUs there is none and the buffer keeps the collective
        buf = parallel.FlatGradBuffer(list(G.parameters()) + list(D.parameters()), peer_exchange=True)
        assert buf.peer is None and parallel.PeerExchange.create(64, torch.device("cpu")) is None
        aliased = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(buf.params, buf.views))
        checks = []
        for it in range(3):
            if it == 1:
                G.zero_grad()                # the torch default drops the .grad tensors: the buffer must re-adopt them
                D.zero_grad()
            else:
                buf.zero()
            loss = D(G(x[lo:hi])).sum() / n * world        # mean over ranks of this == the global mean
            loss.backward()
            sent = buf.allreduce(average=True)
            refG = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 5))
            refD = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 1))
            refG.load_state_dict(G.state_dict()); refD.load_state_dict(D.state_dict())
            (refD(refG(x)).sum() / n).backward()
            err = max((a.grad - b.grad).abs().max().item()
                      for a, b in zip(list(G.parameters()) + list(D.parameters()), list(refG.parameters()) + list(refD.parameters())))
            still = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(buf.params, buf.views))
            checks.append((err, sent, still))
        # one model's span only: the other model's gradients stay local
        buf.zero()
        (D(G(x[lo:hi])).sum() * (rank + 1)).backward()
        before_G = [p.grad.clone() for p in G.parameters()]
        buf.allreduce(average=False, span=buf.span_of(D))
        g_untouched = all(torch.equal(a, p.grad) for a, p in zip(before_G, G.parameters()))
        # off NCCL there is no CTA limit to configure: the helper falls back to the default group
        assert parallel.grad_allreduce_group(4) is None
        q.put((rank, aliased, checks, g_untouched, buf.flat.numel(), buf.span_of(G), buf.span_of(D)))
    finally:
        dist.destroy_process_group()


def test_flat_grad_buffer_world2():
    """gradients live in ONE persistent buffer (the slices are the .grad tensors); one collective
    for generator + critic, no cat, no copy back; survives zero_grad(set_to_none=True)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flatbuf_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    padded = lambda k: (k + 3) // 4 * 4
    total = sum(padded(k) for k in (8 * 16, 16, 16 * 5, 5, 5 * 7, 7, 7, 1))
    for rank, aliased, checks, g_untouched, numel, span_g, span_d in res:
        assert aliased and g_untouched
        assert numel == total and span_g[0] == 0 and span_g[1] == span_d[0] and span_d[1] == total
        for err, sent, still in checks:
            assert err < 1e-6 and sent == total and still
