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


@pytest.mark.parametrize("norm_type", ["L2", "L1"])
def test_chamfer_training_losses_match_bruteforce_torch(cuda_device, norm_type):
    """MaskedChamferLoss / MaskedChamferL1Loss / MaskedAdaptiveL1ChamferLoss against the reference formula restated with a
    brute-force torch.cdist nearest neighbour per patch (chamfer_distance_aux.py:153-246 with K = 1), values and
    gradients w.r.t. the prediction; ragged masks (valid prefix)."""
    from deep3dpointclouddenoising_b200.models.losses import (MaskedAdaptiveL1ChamferLoss, MaskedChamferL1Loss,
                                                               MaskedChamferLoss)
    torch.manual_seed(0)
    B, N = 3, 700
    points = torch.randn(B, N, 3, device=cuda_device) * 0.05
    target = torch.randn(B, N, 3, device=cuda_device) * 0.005
    pred = (torch.randn(B, N, 3, device=cuda_device) * 0.005).requires_grad_(True)
    mask = torch.zeros(B, N, device=cuda_device)
    for b, v in enumerate((700, 512, 33)):
        mask[b, :v] = 1

    def ref_chamfer(pred_, norm):
        cd = 0
        for b in range(B):
            m = mask[b].bool()
            x, y = (points + target)[b, m], (points + pred_)[b, m]
            d = torch.cdist(x.double(), y.double())
            ix, iy = d.argmin(1), d.argmin(0)
            if norm == "L2":
                cx, cy = ((x - y[ix]) ** 2).sum(1), ((y - x[iy]) ** 2).sum(1)
            else:
                cx, cy = (x - y[ix]).abs().sum(1), (y - x[iy]).abs().sum(1)
            cd = cd + cx.mean() + cy.mean()
        return cd / B

    def ref_l1(pred_):
        return ((pred_ - target).abs().mean(2) * mask).sum() / mask.sum()

    cases = [(MaskedChamferLoss(norm_type), lambda p: ref_chamfer(p, norm_type)),
             (MaskedChamferL1Loss(norm_type), lambda p: 0.5 * (ref_l1(p) + ref_chamfer(p, norm_type)))]
    if norm_type == "L1":
        cases += [(MaskedAdaptiveL1ChamferLoss('chamfer'), lambda p: ref_l1(p) + torch.exp(-ref_l1(p)) * ref_chamfer(p, "L1")),
                  (MaskedAdaptiveL1ChamferLoss('L1'), lambda p: ref_chamfer(p, "L1") + torch.exp(-ref_chamfer(p, "L1")) * ref_l1(p))]
    for module, ref in cases:
        got = module(pred, target, mask, points)
        exp = ref(pred)
        torch.testing.assert_close(got, exp, rtol=1e-5, atol=1e-9)
        g_got, = torch.autograd.grad(got, pred)
        g_exp, = torch.autograd.grad(exp, pred)
        torch.testing.assert_close(g_got, g_exp, rtol=1e-4, atol=1e-9)
        assert float(g_got[2, 33:].abs().max()) == 0.0  # padded points receive no gradient
