"""Build-container only (skipped where /root/reference is absent): every committed fixture under tests/golden/ is what
the UNMODIFIED reference produces today -- oracle/make_golden.py's generators are re-run and compared with the committed
arrays.  Float arrays must agree to 1e-6 relative (bit-identical on the host that generated them; BLAS/libm may round
differently elsewhere), integer / bool / string arrays exactly."""
import numpy as np
import pytest
import torch

import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not mounted")

NAMES = ["tables", "kat", "gan133", "stress200", "video36", "camera_ops", "sampler40", "generator", "retarget", "critic",
         "video_critic", "gan_loop"]


@pytest.mark.parametrize("name", NAMES)
def test_committed_fixture_is_what_the_reference_produces(golden, name):
    import make_golden
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        fresh = make_golden.fixtures_table()[name]()
    finally:
        torch.set_num_threads(threads)
    committed = golden(name)
    assert sorted(fresh) == sorted(committed)
    for k, v in fresh.items():
        a, b = np.asarray(v), committed[k]
        assert a.shape == b.shape, (name, k)
        if a.dtype.kind == "f":
            scale = np.maximum(np.abs(b.astype(np.float64)), 1.0)
            assert np.all((np.abs(a.astype(np.float64) - b) / scale <= 1e-6) | (np.isnan(a) & np.isnan(b))), (name, k)
        else:
            assert np.array_equal(a, b), (name, k)
