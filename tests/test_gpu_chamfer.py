"""GPU parity, row f3: exact grid-accelerated nearest neighbours and the Chamfer distance
(ref compute_cd.py:74-75, chamfer_distance_aux.py:154-155,216-246) against a CPU KD-tree / brute force."""
import numpy as np
import pytest
import torch
from scipy.spatial import cKDTree

from deep3dpointclouddenoising_b200 import synthetic

pytestmark = pytest.mark.gpu


def _cpu_chamfer(x, y):
    dx = cKDTree(y).query(x, k=1)[0] ** 2
    dy = cKDTree(x).query(y, k=1)[0] ** 2
    return dx.mean() + dy.mean(), dx, dy


@pytest.mark.parametrize("n,m", [(1, 1), (37, 500), (5000, 4000), (100000, 120000)])
def test_nn_and_chamfer_match_kdtree(cuda_device, n, m):
    from deep3dpointclouddenoising_b200 import ops
    rng = np.random.default_rng(n + m)
    clean = synthetic.make_cloud(1, m, sigma=0.0)
    noisy = clean[rng.integers(0, m, n)] + rng.standard_normal((n, 3)).astype(np.float32) * 0.004
    if n > 30:
        noisy[:7] += 3.0  # far outliers outside the support grid: many empty shells, still exact
    x, y = torch.from_numpy(noisy.astype(np.float32)).to(cuda_device), torch.from_numpy(clean).to(cuda_device)
    d2, idx = ops.nn_sqdist(x, y, want_idx=True)
    ref_d, ref_i = cKDTree(clean.astype(np.float64)).query(noisy.astype(np.float64), k=1)
    np.testing.assert_allclose(d2.cpu().numpy(), ref_d ** 2, rtol=2e-5, atol=1e-12)
    # the index may differ only where two supports are equidistant within float rounding
    got = np.linalg.norm(clean[idx.cpu().numpy()].astype(np.float64) - noisy, axis=1)
    np.testing.assert_allclose(got, ref_d, rtol=1e-5, atol=1e-9)
    cd = ops.chamfer_l2(x, y).cpu().numpy()
    ref, dx, dy = _cpu_chamfer(noisy.astype(np.float64), clean.astype(np.float64))
    np.testing.assert_allclose(cd, [ref, dx.mean(), dy.mean()], rtol=2e-5)
    assert np.array_equal(cd, ops.chamfer_l2(x, y).cpu().numpy())  # deterministic


def test_chamfer_identical_clouds_is_zero_and_symmetric(cuda_device):
    from deep3dpointclouddenoising_b200 import ops
    a = torch.from_numpy(synthetic.make_cloud(2, 30000)).to(cuda_device)
    b = torch.from_numpy(synthetic.make_cloud(3, 20000)).to(cuda_device)
    assert ops.chamfer_l2(a, a)[0].item() == 0.0
    ab, ba = ops.chamfer_l2(a, b).cpu().numpy(), ops.chamfer_l2(b, a).cpu().numpy()
    np.testing.assert_allclose(ab[0], ba[0], rtol=1e-6)
    np.testing.assert_allclose(ab[1], ba[2], rtol=1e-6)
