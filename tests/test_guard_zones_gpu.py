"""Out-of-bounds WRITE check for every kernel family (compute-sanitizer is not available on the GPU pool): every
output lives inside one arena, separated by guard zones holding a sentinel bit pattern; after the launches the
guards must be untouched and every output element must have been written.  Sizes straddle the 32-row tile."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 1024                      # floats on either side of every output
SENT = 0x7FC0DEAD                 # a quiet-NaN payload no kernel produces


class Arena:
    def __init__(self, floats):
        self.buf = torch.full((floats,), 0, dtype=torch.int32, device="cuda:0")
        self.buf.fill_(SENT)
        self.off = 0
        self.outs = []

    def out(self, *shape):
        n = int(np.prod(shape))
        self.off = (self.off + GUARD + 3) // 4 * 4          # 16-byte aligned start after a guard
        t = self.buf[self.off:self.off + n].view(torch.float32).view(*shape)
        self.outs.append((self.off, n))
        self.off += n
        return t

    def check(self):
        torch.cuda.synchronize()
        mask = torch.ones_like(self.buf, dtype=torch.bool)
        for off, n in self.outs:
            mask[off:off + n] = False
            assert not (self.buf[off:off + n] == SENT).any(), "an output element was never written"
        assert (self.buf[mask] == SENT).all(), "a kernel wrote outside its output"


@pytest.mark.parametrize("n", [1, 31, 32, 33, 95, 1000])
def test_no_kernel_writes_outside_its_outputs(n):
    from dhfk import _cabi, synthetic, tables
    lib = _cabi.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    d = {k: torch.tensor(v, device=dev) for k, v in synthetic.gan_like(n, seed=n).items()}
    up = {k: torch.tensor(v, device=dev) for k, v in synthetic.upstream_grads(n, seed=n + 1).items()}
    blk = tables.camera_block("S1", 0)
    A = Arena(64 * GUARD + 4096 * n)
    P = lambda t: t.data_ptr()
    world, cam, uv = A.out(n, 16, 3), A.out(n, 16, 3), A.out(n, 16, 2)
    for flags in (0, _cabi.FLAG_FAST_TRIG):
        _cabi.check(lib.dhfk_forward(P(d["ang"]), 33, P(d["grot"]), 3, P(d["bone"]), 15, P(d["root"]), 3, blk.ctypes.data,
                                     P(world), P(cam), P(uv), n, flags, st), "fwd")
    g_ang, g_grot, g_root, g_bone = A.out(n, 33), A.out(n, 3), A.out(n, 3), A.out(n, 15)
    for flags in (0, _cabi.FLAG_ACCURATE_TRIG):
        _cabi.check(lib.dhfk_backward(P(d["ang"]), 33, P(d["grot"]), 3, P(d["bone"]), 15, P(d["root"]), 3, blk.ctypes.data,
                                      P(up["g_world"]), P(up["g_cam"]), P(up["g_uv"]), P(g_ang), 33, P(g_grot), 3,
                                      P(g_root), 3, P(g_bone), 15, n, flags, st), "bwd")
    # strided gradient outputs (a [n,37] angle-gradient view): only columns 0..32 of each row may be written
    g37 = A.out(n, 37)
    g37.fill_(7.0)
    _cabi.check(lib.dhfk_backward(P(d["ang"]), 33, P(d["grot"]), 3, P(d["bone"]), 15, P(d["root"]), 3, blk.ctypes.data,
                                  P(up["g_world"]), None, P(up["g_uv"]), P(g37), 37, P(g_grot), 3, P(g_root), 3,
                                  None, 15, n, 0, st), "bwd strided")
    torch.cuda.synchronize()
    assert (g37[:, 33:] == 7.0).all() and not (g37[:, :33] == 7.0).all()
    # generator mode
    half, mid = tables.generator_slot_scale(True)
    raw = torch.randn(n, 35, device=dev)
    gworld, guv, d_raw = A.out(n, 16, 3), A.out(n, 16, 2), A.out(n, 35)
    _cabi.check(lib.dhfk_generator_forward(P(raw), 35, P(d["bone"]), 15, half.ctypes.data, mid.ctypes.data, 10.0,
                                           blk.ctypes.data, P(gworld), None, P(guv), n, 0, st), "gen fwd")
    _cabi.check(lib.dhfk_generator_backward(P(raw), 35, P(d["bone"]), 15, half.ctypes.data, mid.ctypes.data, 10.0,
                                            blk.ctypes.data, P(up["g_world"]), None, P(up["g_uv"]), P(d_raw), 35, n, 0, st),
                "gen bwd")
    # retarget + projection
    pose = torch.randn(n, 16, 3, device=dev) + torch.tensor([0.0, 0.0, 5.0], device=dev)
    idx = torch.randint(0, 5, (n,), device=dev, dtype=torch.int32)
    tm = torch.tensor(tables.BONE_TEMPLATES_GANUTILS_ORDER, device=dev)
    rows = torch.tensor(blk[7:16], device=dev).repeat(n, 1).contiguous()
    rp, ruv = A.out(n, 16, 3), A.out(n, 16, 2)
    _cabi.check(lib.dhfk_retarget_project(P(pose), P(idx), P(tm), 5, P(rows), 9, P(rp), P(ruv), n, st), "retarget")
    # critic inputs: forward (30 and 15 columns), vjp, jvp, flips
    cpos, k30, k15 = A.out(n, 16, 3), A.out(n, 30), A.out(n, 15)
    _cabi.check(lib.dhfk_critic_input_forward(P(pose), P(cpos), P(k30), 30, n, 3, st), "critic fwd")
    _cabi.check(lib.dhfk_critic_input_forward(P(pose), None, P(k15), 15, n, 0, st), "critic fwd 15")
    gk = torch.randn(n, 30, device=dev)
    gpose, tpos, tk = A.out(n, 16, 3), A.out(n, 16, 3), A.out(n, 30)
    _cabi.check(lib.dhfk_critic_input_backward(P(pose), P(up["g_world"]), P(gk), 30, P(gpose), n, 1, st), "critic vjp")
    _cabi.check(lib.dhfk_critic_input_jvp(P(pose), P(up["g_world"]), P(tpos), P(tk), 30, n, 1, st), "critic jvp")
    f3, f2 = A.out(n, 16, 3), A.out(n, 16, 2)
    _cabi.check(lib.dhfk_flip_pose(P(pose), P(f3), n, 3, st), "flip3")
    _cabi.check(lib.dhfk_flip_pose(P(up["g_uv"]), P(f2), n, 2, st), "flip2")
    # standalone camera ops
    w2c, puv, gx = A.out(n, 16, 3), A.out(n, 16, 2), A.out(n, 16, 3)
    q, t = torch.tensor(blk[0:4], device=dev), torch.tensor(blk[4:7], device=dev)
    _cabi.check(lib.dhfk_world_to_camera_forward(P(world), P(q), P(t), 1, P(w2c), n * 16, st), "w2c")
    _cabi.check(lib.dhfk_project_forward(P(pose), P(rows), 9, P(puv), n, 16, st), "project")
    _cabi.check(lib.dhfk_project_backward(P(pose), P(rows), 9, P(up["g_uv"]), P(gx), n, 16, st), "project bwd")
    # bank gather
    rec = torch.randn(max(n, 4), 96, device=dev)
    perm = torch.randint(0, rec.shape[0], (n,), device=dev)
    b3, b2, bc = A.out(n, 16, 3), A.out(n, 16, 2), A.out(n, 9)
    _cabi.check(lib.dhfk_bank_gather(P(rec), 96, 9, P(perm), n, rec.shape[0], P(b3), P(b2), P(bc), st), "bank")
    A.check()
