"""BASELINE configs[2] / configs[3] with the reference's OWN training loops: the unmodified
GAN_solutions_FK_generator (models_Fk_GAN/model_fk_gan_train.py:236-512) and
video_mode_GAN_solutions_FK_generator (models_Fk_GAN/video_GAN_fun.py:79-602), imported from the archive
oracle/stage_ref.py staged, run on the GPU with dhfk.dropin.install(...) and compared, iteration by iteration, with the
same functions run unpatched on the CPU in the build container (tests/golden/gan_loop.npz, oracle/ref_loop.py).

What is compared: every scalar the loop reports (D_real, D_fake, Wasserstein distance of every critic step), the
generator's gradients at its step, the fake-pair buffer (pos_3d_cam / uv of every iteration).  The MLPs run on cuBLAS
here and on the CPU GEMM there and five Adam steps lie between the first and the last number, so the bound is looser
than the kernels' own 1e-5 (which tests/test_parity_gpu.py etc. hold them to): 2e-3 absolute on O(0.1) losses."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_archive():
    import ref_harness as rh
    assert rh.staged_reference_available() or rh.reference_available(), \
        "the staged reference archive is missing: run `python oracle/stage_ref.py` in the build container " \
        "(it is git-ignored and travels to the GPU box with the snapshot)"


def _compare(r, g, mode):
    import ref_loop
    names, vals = ref_loop.scalars_matrix(r["scalars"])
    assert names == [str(x) for x in g[mode + "_scalar_names"]], "the loop reported a different sequence of scalars"
    ref = g[mode + "_scalars"]
    err = np.abs(vals - ref)
    worst = int(err.argmax())
    assert err.max() <= 2e-3, "%s: scalar %d (%s) differs by %.3e (got %.6f, reference %.6f)" % (
        mode, worst, names[worst], err.max(), vals[worst], ref[worst])
    gg, gref = np.stack(r["g_grads"]), g[mode + "_g_grads"]
    assert gg.shape == gref.shape
    scale = max(1e-6, float(np.abs(gref).max()))
    assert np.abs(gg - gref).max() <= 2e-2 * scale, "%s: generator gradient differs by %.3e of its max" % (
        mode, np.abs(gg - gref).max() / scale)
    for k in ("buffer_3d", "buffer_2d"):
        got, want = np.asarray(r[k], np.float32), g["%s_%s" % (mode, k)]
        assert got.shape == want.shape
        # fakes come out of a generator whose weights went through Adam steps on both sides: positions agree to ~1e-3 m
        assert np.abs(got - want).max() <= 5e-3, "%s: %s differs by %.3e" % (mode, k, np.abs(got - want).max())
    return float(err.max()), float(np.abs(gg - gref).max() / scale)


@pytest.mark.parametrize("mode,install", [
    ("single", dict()),                                                   # FK class + camera functions only
    ("single", dict(generators=True, critics=True, loader_refresh=True)),
    ("video", dict()),
    ("video", dict(generators=True, critics=True, loader_refresh=True)),
])
def test_reference_loop_with_dropin_matches_unpatched_cpu_run(golden, mode, install):
    _need_archive()
    import ref_loop
    g = golden("gan_loop")
    iters, batch, dense, seed = (int(x) for x in g[mode + "_cfg"])
    r = ref_loop.run_loop(mode, device="cuda", iters=iters, batch=batch, dense=dense, seed=seed, install=install)
    e_s, e_g = _compare(r, g, mode)
    print("[ref loop %s %s] max scalar diff %.2e, generator-gradient diff %.2e of max" % (mode, sorted(install), e_s, e_g))
    import torch
    mods = ref_loop.load(force_cpu=False)
    import dhfk
    # the loop really ran on the native classes / functions
    assert sys.modules["models_Fk_GAN.forward_kinematics_DH_model"].Forward_Kinematics_DH_Model is dhfk.Forward_Kinematics_DH_Model
    assert mods["train"].project_to_2d is dhfk.camera.project_to_2d
    if install.get("critics"):
        assert mods["dis"].Video_motion_Fk_3D_Discriminator.forward is dhfk.Fk_discriminator.video_motion_3d_forward
    assert torch.cuda.is_available()


def test_reference_wrap_helper_works_with_the_dropin_installed():
    """traditional_solutions_FK_generator (model_fk_gan_train.py:36-94, `--data_enhancement_method normal`) calls
    wrap(project_to_2d, True, numpy, numpy): CPU float64 in, `.numpy()` on the result.  With the drop-in installed that
    call must still work (it used to get a CUDA tensor back)."""
    _need_archive()
    import ref_loop
    import dhfk
    mods = ref_loop.load(force_cpu=False)
    dhfk.dropin.install()
    from utils.utils import wrap                      # the reference's own helper (utils/utils.py:137-165)
    from dhfk import tables
    rng = np.random.RandomState(0)
    pose = rng.randn(50, 16, 3) * 0.3 + np.array([0.0, 0.0, 4.5])
    cam9 = tables.camera_block("S1", 0)[7:16].astype(np.float64)
    out = wrap(mods["train"].project_to_2d, True, pose, cam9)
    assert isinstance(out, np.ndarray) and out.shape == (50, 16, 2)
    import torch_port
    import torch
    want = torch_port.project_to_2d(torch.tensor(pose, dtype=torch.float32),
                                    torch.tensor(cam9, dtype=torch.float32).view(1, 9).repeat(50, 1)).numpy()
    assert np.abs(out - want).max() <= 1e-5
