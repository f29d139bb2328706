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
