"""Parity of the sm_100a kernels against (a) the golden vectors produced by the unmodified reference
and (b) the float64 C oracle on seeded inputs, through the public API (-> ctypes -> C ABI).

Tolerance (BASELINE.json north_star): |x - ref| <= 1e-5 * max(|ref|, 1), fp32.
Gradients that pass through x/z use the conditioning multiplier of conftest.projection_conditioning
for poses with a joint closer than 0.5 m to the camera plane (multiplier 1 everywhere else)."""
import numpy as np
import pytest
import torch

from conftest import RTOL, assert_parity, assert_parity_vs_reference, projection_conditioning, rel_err

pytestmark = pytest.mark.gpu

CASES = ["gan133", "stress200", "video36"]


def dev():
    return torch.device("cuda", 0)


def T(x, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float32), device=dev(), requires_grad=grad)


TRIG = {"accurate": dict(accurate_grad=True), "default": {}, "mufu": dict(fast_trig=True)}   # include/dhfk.h flags
TRIG_IDS = list(TRIG)


def run_fused(g, trig, grads):
    """grads: subset string of 'wcu' -> returns outputs and input gradients as numpy."""
    import dhfk
    ang, grot, root = T(g["ang"], True), T(g["grot"], True), T(g["root"], True)
    bone = T(g["bone"])
    world, cam, uv = dhfk.fk_project(ang, grot, bone, root, g["cam_block"], return_cam=True, **TRIG[trig])
    loss = (world * T(g["g_world"])).sum()
    if "c" in grads:
        loss = loss + (cam * T(g["g_cam"])).sum()
    if "u" in grads:
        loss = loss + (uv * T(g["g_uv"])).sum()
    loss.backward()
    return (world.detach().cpu().numpy(), cam.detach().cpu().numpy(), uv.detach().cpu().numpy(),
            ang.grad.cpu().numpy(), grot.grad.cpu().numpy(), root.grad.cpu().numpy())


@pytest.mark.parametrize("trig", TRIG_IDS)
@pytest.mark.parametrize("case", CASES)
def test_forward_matches_reference_golden(golden, case, trig):
    g = golden(case)
    world, cam, uv, *_ = run_fused(g, trig, "w")
    assert_parity(world, g["world16"], "world16")
    assert_parity(cam, g["cam"], "cam")
    assert_parity(uv, g["uv"], "uv")


@pytest.mark.parametrize("trig", TRIG_IDS)
@pytest.mark.parametrize("tag", ["w", "wu", "wcu"])
@pytest.mark.parametrize("case", CASES)
def test_backward_matches_reference_autograd(golden, c_oracle, case, tag, trig):
    """Against the reference's own autograd (goldens).  The bound is reference-relative (conftest.
    assert_parity_vs_reference): the plain 1e-5 for every pose the reference itself computes accurately -- all of gan133 and
    video36 -- plus, for the clamp-active / near-camera-plane poses of stress200, 4x the reference's OWN fp32-vs-float64
    error on that pose."""
    g = golden(case)
    *_, g_ang, g_grot, g_root = run_fused(g, trig, tag)
    b = c_oracle.backward(g["ang"], g["grot"], g["bone"], g["root"], g["cam_block"], g_world=g["g_world"],
                          g_cam=g["g_cam"] if "c" in tag else None, g_uv=g["g_uv"] if "u" in tag else None, want_bone=False)
    needed = 0
    for name, x in (("g_ang", g_ang), ("g_grot", g_grot), ("g_root", g_root)):
        _, k = assert_parity_vs_reference(x, g["%s_%s" % (name, tag)], b[name], "%s %s %s" % (case, tag, name))
        needed += k
    if case != "stress200":
        assert needed == 0, "a normally placed pose needed the reference-error allowance"
    assert np.all(g_ang[:, [4, 9, 22, 27, 32]] == 0)


def test_kat_tpose_and_bent(golden):
    import dhfk
    g = golden("kat")
    bone = np.array([[.5, .5, .6, .6, .25, .25, .25, .2, .4, .4, .4, .4, .35, .35, .15]], np.float32)
    w = dhfk.fk_world16(T(np.zeros((1, 33))), T(np.zeros((1, 3))), T(bone), T(np.zeros((1, 3))))
    idx = [0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27]
    assert np.abs(w.cpu().numpy()[0] - g["tpose32"][idx]).max() < 1e-6
    for pre in ("bent_", "bent2_"):
        gg = {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}
        world, cam, uv, *_ = run_fused(gg, "default", "w")
        assert_parity(world, gg["world16"], pre + "world16")
        assert_parity(uv, gg["uv"], pre + "uv")


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 95, 96, 97, 1000, 4608])
def test_ragged_sizes_vs_c_oracle(c_oracle, n):
    """Tile edges (32 rows per warp/CTA), single pose, BASELINE cfg-4 size 512*9."""
    import dhfk
    from dhfk import synthetic, tables
    inp = synthetic.gan_like(n, seed=100 + n)
    up = synthetic.upstream_grads(n, seed=7 + n)
    blk = tables.camera_block("S5", 1)
    g = dict(inp, cam_block=blk, **up)
    world, cam, uv, g_ang, g_grot, g_root = run_fused(g, "default", "wcu")
    o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk)
    b = c_oracle.backward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk, g_world=up["g_world"],
                          g_cam=up["g_cam"], g_uv=up["g_uv"])
    assert_parity(world, o["world16"], "world16"); assert_parity(cam, o["cam"], "cam"); assert_parity(uv, o["uv"], "uv")
    assert_parity(g_ang, b["g_ang"], "g_ang"); assert_parity(g_grot, b["g_grot"], "g_grot")
    assert_parity(g_root, b["g_root"], "g_root")


def test_every_training_camera_vs_c_oracle(c_oracle):
    """the 20 (train subject, camera) pairs the GAN loop draws from (model_fk_gan_train.py:344-372)."""
    from dhfk import synthetic, tables
    n = 224
    for si, subj in enumerate(tables.TRAIN_SUBJECTS):
        for cam_id in range(4):
            inp = synthetic.gan_like(n, seed=900 + 4 * si + cam_id)
            up = synthetic.upstream_grads(n, seed=950 + 4 * si + cam_id)
            blk = tables.camera_block(subj, cam_id)
            g = dict(inp, cam_block=blk, **up)
            world, cam, uv, g_ang, g_grot, g_root = run_fused(g, "default", "wcu")
            o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk)
            b = c_oracle.backward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk, g_world=up["g_world"],
                                  g_cam=up["g_cam"], g_uv=up["g_uv"])
            tag = " %s cam %d" % (subj, cam_id)
            assert_parity(world, o["world16"], "world16" + tag); assert_parity(cam, o["cam"], "cam" + tag)
            assert_parity(uv, o["uv"], "uv" + tag)
            assert_parity(g_ang, b["g_ang"], "g_ang" + tag); assert_parity(g_grot, b["g_grot"], "g_grot" + tag)
            assert_parity(g_root, b["g_root"], "g_root" + tag)


def test_empty_batch():
    import dhfk
    w, c, u = dhfk.fk_project(T(np.zeros((0, 33))), T(np.zeros((0, 3))), T(np.zeros((0, 15))), T(np.zeros((0, 3))),
                              np.zeros(16, np.float32))
    assert w.shape == (0, 16, 3) and c.shape == (0, 16, 3) and u.shape == (0, 16, 2)


def test_strided_views_and_bone_gradient(c_oracle):
    """The generator hands FK column slices of one [N,37] tensor and root as a slice of [N,35]
    (Fk_generator.py:126,179-186): non-packed rows take the gather path.  Also checks d/d(bone)."""
    import dhfk
    from dhfk import synthetic, tables
    n = 333
    inp = synthetic.gan_like(n, seed=21)
    up = synthetic.upstream_grads(n, seed=22)
    blk = tables.camera_block("S6", 2)
    gen = torch.zeros(n, 37, device=dev())
    gen[:, :33] = T(inp["ang"]); gen[:, 34:] = T(inp["grot"])
    gen.requires_grad_(True)
    netout = torch.zeros(n, 35, device=dev()); netout[:, -3:] = T(inp["root"]); netout.requires_grad_(True)
    bone = T(inp["bone"], True)
    world, cam, uv = dhfk.fk_project(gen[:, 0:33], gen[:, -3:], bone, netout[:, -3:], blk)
    ((world * T(up["g_world"])).sum() + (uv * T(up["g_uv"])).sum()).backward()
    o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk)
    b = c_oracle.backward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk, g_world=up["g_world"], g_uv=up["g_uv"])
    assert_parity(world.detach().cpu().numpy(), o["world16"], "world16")
    assert_parity(uv.detach().cpu().numpy(), o["uv"], "uv")
    gg = gen.grad.cpu().numpy()
    assert_parity(gg[:, :33], b["g_ang"], "g_ang"); assert_parity(gg[:, 34:], b["g_grot"], "g_grot")
    assert np.all(gg[:, 33] == 0)
    assert_parity(netout.grad.cpu().numpy()[:, -3:], b["g_root"], "g_root")
    assert_parity(bone.grad.cpu().numpy(), b["g_bone"], "g_bone")
    # full [N,37] tensor passed directly (stride 37, 33 columns used)
    w2 = dhfk.fk_world16(gen.detach(), gen.detach()[:, -3:], bone.detach(), netout.detach()[:, -3:])
    assert torch.equal(w2, world.detach())


def test_fk_only_and_no_cam_variants_agree():
    import dhfk
    from dhfk import synthetic, tables
    inp = synthetic.gan_like(500, seed=5)
    a, g, b, r = (T(inp[k]) for k in ("ang", "grot", "bone", "root"))
    blk = tables.camera_block("S1", 0)
    w0 = dhfk.fk_world16(a, g, b, r)
    w1, c1, u1 = dhfk.fk_project(a, g, b, r, blk, return_cam=True)
    w2, c2, u2 = dhfk.fk_project(a, g, b, r, blk, return_cam=False)
    assert c2 is None and torch.equal(w0, w1) and torch.equal(w0, w2) and torch.equal(u1, u2)
    # skeleton invariant: every bone keeps its length whatever the angles are
    from dhfk.tables import used_16key_15bone_len_table
    w = w0.cpu().numpy().astype(np.float64)
    for bi, (i, j) in enumerate(used_16key_15bone_len_table):
        L = np.linalg.norm(w[:, i] - w[:, j], axis=1)
        assert np.abs(L - inp["bone"][:, bi]).max() < 2e-6, bi


@pytest.mark.parametrize("trig", TRIG_IDS)
def test_full_size_1m_vs_c_oracle(c_oracle, trig):
    """BASELINE config 2 size (1,048,576 poses): forward and backward against the float64 oracle on
    every pose, plus size-independent properties (linearity of the backward in the upstream gradient,
    root-translation equivariance)."""
    import dhfk
    from dhfk import synthetic, tables
    n = 1 << 20
    inp = synthetic.gan_like(n, seed=1234)
    up = synthetic.upstream_grads(n, seed=4321)
    blk = tables.camera_block("S1", 0)
    g = dict(inp, cam_block=blk, **up)
    world, cam, uv, g_ang, g_grot, g_root = run_fused(g, trig, "wu")
    o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk)
    assert (np.abs(o["cam"][..., :2] / o["cam"][..., 2:]) < 1).all()     # in-volume roots: clamp inactive
    e = [assert_parity(world, o["world16"], "world16"), assert_parity(cam, o["cam"], "cam"),
         assert_parity(uv, o["uv"], "uv")]
    b = c_oracle.backward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk, g_world=up["g_world"],
                          g_uv=up["g_uv"], want_bone=False)
    e += [assert_parity(g_ang, b["g_ang"], "g_ang"), assert_parity(g_grot, b["g_grot"], "g_grot"),
          assert_parity(g_root, b["g_root"], "g_root")]
    print("\n[1M %s] max rel err world/cam/uv/g_ang/g_grot/g_root = %s" % (trig,
                                                                          " ".join("%.2e" % x for x in e)))
    # linearity: backward(2*g) == 2*backward(g) bit-exactly (power-of-two scaling commutes with fp32 rounding)
    g2 = dict(g, g_world=2 * up["g_world"], g_uv=2 * up["g_uv"])
    *_, a2, r2, t2 = run_fused(g2, trig, "wu")
    assert np.array_equal(a2, 2 * g_ang) and np.array_equal(r2, 2 * g_grot) and np.array_equal(t2, 2 * g_root)


@pytest.mark.parametrize("trig", TRIG_IDS)
def test_stress_1m_clamp_active(c_oracle, trig):
    """1M poses with roots 10*tanh(randn): most poses leave the image (clamp active) or sit behind the
    camera.  Forward must match everywhere.  Gradients are compared with the conditioning multiplier;
    poses with a point whose x/z sits on the clamp edge within fp32 resolution are excluded, because
    torch.clamp's gradient mask is discontinuous there (an fp32 reference and a float64 oracle can
    legitimately disagree about which side the point is on)."""
    import dhfk
    from dhfk import synthetic, tables
    n = 1 << 20
    inp = synthetic.gan_like(n, seed=77, root_mode="generator", angle_mode="stress")
    up = synthetic.upstream_grads(n, seed=78)
    blk = tables.camera_block("S8", 3)
    g = dict(inp, cam_block=blk, **up)
    world, cam, uv, g_ang, g_grot, g_root = run_fused(g, trig, "wu")
    o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk)
    assert_parity(world, o["world16"], "world16"); assert_parity(cam, o["cam"], "cam")
    ratio = np.abs(o["cam"][..., :2] / o["cam"][..., 2:])
    assert (ratio > 1).mean() > 0.2
    # fp32 resolution of x/z: camera coordinates (|X| <= ~25 m) carry ~2e-6 m absolute rounding error
    band = 4e-6 * (1.0 + ratio) / np.abs(o["cam"][..., 2:])
    edge = (np.abs(ratio - 1) < band).any(axis=(1, 2))
    cond = projection_conditioning(o["cam"], o["world16"], blk, up["g_uv"])
    ok = ~edge
    assert ok.mean() > 0.999
    assert_parity(uv[ok], o["uv"][ok], "uv", row_scale=projection_conditioning(o["cam"], o["world16"], blk, kind="uv")[ok])
    b = c_oracle.backward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk, g_world=up["g_world"],
                          g_uv=up["g_uv"], want_bone=False)
    for name, x, ref in (("g_ang", g_ang, b["g_ang"]), ("g_grot", g_grot, b["g_grot"]), ("g_root", g_root, b["g_root"])):
        err = (np.abs(x.astype(np.float64) - ref) / np.maximum(np.abs(ref), 1.0)).max(axis=1) / cond
        err[~ok] = 0
        worst = int(np.argmax(err))
        assert np.isfinite(err).all() and err[worst] <= RTOL, (
            "%s: worst pose %d scaled err %.3e (cond %.1f, min|z| %.4g, ratios %s)" % (
                name, worst, err[worst], cond[worst], np.abs(o["cam"][worst, :, 2]).min(),
                np.round(ratio[worst].ravel(), 6).tolist()))


def test_host_pipeline_matches_device_path():
    """dhfk_forward_backward_host (pinned host buffers, chunked copies) == device-resident path, bit-exact."""
    import dhfk
    from dhfk import synthetic, tables
    n = 50000
    inp = synthetic.gan_like(n, seed=9)
    up = synthetic.upstream_grads(n, seed=10)
    blk = tables.camera_block("S7", 0)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    res = dhfk.fk_project_host(pin(inp["ang"]), pin(inp["grot"]), pin(inp["bone"]), pin(inp["root"]), blk,
                               pin(up["g_world"]), pin(up["g_uv"]), chunk_rows=8192, num_streams=3)
    g = dict(inp, cam_block=blk, **up)
    world, cam, uv, g_ang, g_grot, g_root = run_fused(g, "default", "wu")
    assert np.array_equal(res["world"].numpy(), world) and np.array_equal(res["uv"].numpy(), uv)
    assert np.array_equal(res["g_ang"].numpy(), g_ang) and np.array_equal(res["g_grot"].numpy(), g_grot)
    assert np.array_equal(res["g_root"].numpy(), g_root)


def test_multi_million_batch_crosses_2gib_offsets(c_oracle):
    """BASELINE config 5 shards 16 M poses over 2-8 GPUs (2-8 M per GPU).  One launch over 11.3 M poses puts byte
    offsets of world16 beyond 2^31 and element offsets of the angle tensor beyond 2^28: spot-check windows at
    the start, the end and around the 2^31-byte offsets against the float64 oracle (64-bit indexing everywhere)."""
    import dhfk
    from dhfk import synthetic, tables
    n = (1 << 31) // 192 + 150_001
    d = synthetic.gan_like_torch(n, dev(), seed=3)
    blk = tables.camera_block("S5", 2)
    a, g, r = d["ang"].requires_grad_(True), d["grot"].requires_grad_(True), d["root"].requires_grad_(True)
    w, _, uv = dhfk.fk_project(a, g, d["bone"], r, blk, return_cam=False)
    gen = torch.Generator(device=dev()).manual_seed(5)
    gw = torch.randn((n, 16, 3), device=dev(), generator=gen)
    gu = torch.randn((n, 16, 2), device=dev(), generator=gen)
    ((w * gw).sum() + (uv * gu).sum()).backward()
    for lo in (0, n - 2048, (1 << 31) // 192 - 1024, (1 << 31) // 192 + 100_000, (1 << 30) // 132 - 1024):
        sl = slice(lo, lo + 2048)
        c = lambda t: t[sl].detach().cpu().numpy()
        o = c_oracle.forward(c(a), c(g), c(d["bone"]), c(r), blk)
        b = c_oracle.backward(c(a), c(g), c(d["bone"]), c(r), blk, g_world=c(gw), g_uv=c(gu), want_bone=False)
        assert_parity(c(w), o["world16"], "world16 @%d" % lo)
        assert_parity(c(uv), o["uv"], "uv @%d" % lo)
        assert_parity(c(a.grad), b["g_ang"], "g_ang @%d" % lo)
        assert_parity(c(g.grad), b["g_grot"], "g_grot @%d" % lo)
        assert_parity(c(r.grad), b["g_root"], "g_root @%d" % lo)
    del a, g, r, w, uv, gw, gu, d
    torch.cuda.empty_cache()


def test_cuda_graph_capture_of_forward_and_backward():
    """The C-ABI launches are plain stream work: torch.cuda.graph captures the fused forward + backward, and replays
    on new data (copied into the static input tensors) reproduce the eagerly launched results bit for bit."""
    import dhfk
    from dhfk import synthetic, tables
    n = 4096 + 13
    blk = tables.camera_block("S6", 1)
    a0, a1 = synthetic.gan_like(n, seed=1), synthetic.gan_like(n, seed=2)
    up = synthetic.upstream_grads(n, seed=3)
    gw, gu = T(up["g_world"]), T(up["g_uv"])
    ang, grot, root, bone = T(a0["ang"], True), T(a0["grot"], True), T(a0["root"], True), T(a0["bone"])

    def step():
        for t in (ang, grot, root):
            t.grad = None
        w, _, uv = dhfk.fk_project(ang, grot, bone, root, blk, return_cam=False)
        ((w * gw).sum() + (uv * gu).sum()).backward()
        return w, uv

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        w_s, uv_s = step()
    g_s = (ang.grad, grot.grad, root.grad)
    for inp in (a1, a0):
        with torch.no_grad():
            ang.copy_(T(inp["ang"])); grot.copy_(T(inp["grot"])); root.copy_(T(inp["root"])); bone.copy_(T(inp["bone"]))
        graph.replay()
        torch.cuda.synchronize()
        got = [t.clone() for t in (w_s, uv_s) + g_s]
        e_ang, e_grot, e_root = T(inp["ang"], True), T(inp["grot"], True), T(inp["root"], True)
        w, _, uv = dhfk.fk_project(e_ang, e_grot, T(inp["bone"]), e_root, blk, return_cam=False)
        ((w * gw).sum() + (uv * gu).sum()).backward()
        for x, ref in zip(got, (w, uv, e_ang.grad, e_grot.grad, e_root.grad)):
            assert torch.equal(x, ref.detach())


@pytest.mark.parametrize("n,chunk,slots,bwd", [(1, 4096, 1, True), (5000, 4096, 1, True), (5000, 4096, 2, False),
                                               (20000, 4096, 3, True), (4096, 4096, 8, True)])
def test_host_pipeline_shapes(n, chunk, slots, bwd):
    """dhfk_forward_backward_host: single slot, ragged last chunk, fewer rows than one chunk, forward-only mode and
    more slots than chunks all give the device path's results bit for bit."""
    import dhfk
    from dhfk import synthetic, tables
    inp = synthetic.gan_like(n, seed=n)
    up = synthetic.upstream_grads(n, seed=n + 1)
    blk = tables.camera_block("S5", 3)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    res = dhfk.fk_project_host(pin(inp["ang"]), pin(inp["grot"]), pin(inp["bone"]), pin(inp["root"]), blk,
                               pin(up["g_world"]) if bwd else None, pin(up["g_uv"]) if bwd else None,
                               chunk_rows=chunk, num_streams=slots)
    g = dict(inp, cam_block=blk, **up)
    world, cam, uv, g_ang, g_grot, g_root = run_fused(g, "default", "wu")
    assert np.array_equal(res["world"].numpy(), world) and np.array_equal(res["uv"].numpy(), uv)
    if bwd:
        assert np.array_equal(res["g_ang"].numpy(), g_ang) and np.array_equal(res["g_grot"].numpy(), g_grot)
        assert np.array_equal(res["g_root"].numpy(), g_root)
    else:
        assert "g_ang" not in res
    with pytest.raises(ValueError):
        dhfk.fk_project_host(pin(inp["ang"]), pin(inp["grot"]), pin(inp["bone"]), pin(inp["root"]), blk,
                             pin(up["g_world"]), None)


def test_projection_nan_and_zero_depth_semantics_match_torch_clamp():
    """common/camera.py:85: XX = torch.clamp(X[..., :2] / X[..., 2:], -1, 1).  torch.clamp propagates NaN (0/0 at
    x = z = 0, NaN inputs) and x/0 = +-inf clamps to +-1; the kernels use FMNMX.NAN for the same result.  Checked against
    the reference's op sequence in torch (oracle/torch_port.py::project_to_2d) on the CPU, forward and gradient:
    identical NaN pattern, finite values within the north-star tolerance.  The one documented divergence
    (include/dhfk.h): a DENORMAL z is flushed to zero by MUFU.RCP, so (x = 0, |z| < 1.2e-38) yields NaN instead of 0."""
    import dhfk
    import torch_port
    nan, inf = float("nan"), float("inf")
    pts = np.array([
        [0.3, -0.2, 4.0], [1.0, 2.0, 0.0], [-1.0, 0.5, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [nan, 0.1, 3.0],
        [0.1, nan, 3.0], [0.1, 0.2, nan], [inf, 0.1, 2.0], [0.1, 0.2, inf], [5.0, -7.0, 1e-30], [0.2, 0.1, -3.0],
        [3.0, 3.0, 3.0], [-3.0, 2.9999, 3.0], [1e-20, 1e-20, 1e-20], [0.5, 0.5, 0.5],
    ], np.float32)
    n = 37
    x = np.tile(pts[None], (n, 1, 1)).copy()
    x[1:] += np.where(np.isfinite(x[1:]) & (x[1:] != 0), np.random.RandomState(3).randn(n - 1, 16, 3).astype(np.float32) * 1e-3, 0)
    rows = np.tile(tables_cam_rows(), (n, 1))
    g_uv = np.random.RandomState(4).randn(n, 16, 2).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True)
    ref = torch_port.project_to_2d(xt, torch.tensor(rows))
    (ref * torch.tensor(g_uv)).sum().backward()
    xd = T(x, True)
    uv = dhfk.project_to_2d(xd, T(rows))
    (uv * T(g_uv)).sum().backward()
    got, want = uv.detach().cpu().numpy(), ref.detach().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)), "uv: NaN pattern differs from torch.clamp semantics"
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin), "uv: inf pattern differs"
    assert (np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1.0)).max() <= 1e-5
    assert np.isnan(want[:, 3]).all() and np.isfinite(want[:, 1]).all()      # 0/0 -> NaN, x/0 -> the clamp edge
    # gradients: wherever torch's own backward stays finite the kernel agrees; a NaN the kernel produces is one torch
    # produces too (torch additionally turns 0 * (x / z^2 = inf) into NaN where the kernel's mask has already zeroed it)
    gg, gw = xd.grad.cpu().numpy(), xt.grad.numpy()
    both = np.isfinite(gw) & np.isfinite(gg)
    assert (np.abs(gg[both] - gw[both]) / np.maximum(np.abs(gw[both]), 1.0)).max() <= 1e-5
    assert not (np.isnan(gg) & np.isfinite(gw)).any(), "the kernel produced a NaN gradient where torch has a finite one"
    for row in (0, 11, 12, 13, 14, 15):
        assert np.isfinite(gg[:, row]).all() and np.isfinite(gw[:, row]).all()


def tables_cam_rows():
    from dhfk import tables
    return tables.camera_block("S1", 0)[7:16].reshape(1, 9).astype(np.float32)
