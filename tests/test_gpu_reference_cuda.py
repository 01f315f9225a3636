"""GPU arbiter: our kernels against the REFERENCE'S OWN CUDA kernels running on the same B200
(oracle/_ref/libref_cuda.so: reference sources compiled by nvcc for sm_100a, see oracle/Makefile).
Bit-exact for every index op and the gather; tolerance for the atomicAdd gradient.  Skipped when the prebuilt
library did not travel with the snapshot (it needs /root/reference at build time)."""
import numpy as np
import pytest
import torch

from deep3dpointclouddenoising_b200 import synthetic
from oracle import cuda_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not cuda_ref.available(), reason="oracle/_ref/libref_cuda.so not present")]


@pytest.fixture(scope="module")
def ref():
    return cuda_ref.RefCuda()


@pytest.mark.parametrize("seed,B,N", [(0, 4, 2048), (1, 2, 777), (2, 16, 8192)])
def test_neighbourhood_pyramid_bit_exact_with_reference_cuda(cuda_device, ref, seed, B, N):
    from deep3dpointclouddenoising_b200 import ops
    pts, mask, _, _ = synthetic.make_batch(900 + seed, B, N, ragged=True)
    xyz, m = torch.from_numpy(pts).to(cuda_device), torch.from_numpy(mask).to(cuda_device)
    dl, radius = 0.003125, 0.025
    nsamples = (52, 39, 32, 26, 26)
    for level in range(3 if N >= 2048 else 2):
        npoint = max(xyz.shape[1] // 4, 1)
        a, b = ops.grid_subsample(xyz, m, npoint, dl), ref.grid_subsampling(xyz, m, npoint, dl)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), f"grid subsample level {level}"
        sub, subm = a
        for q, qm in ((xyz, m), (sub, subm)):
            x, y = ops.ball_query(q, xyz, qm, m, radius, nsamples[level]), ref.ball_query(q, xyz, qm, m, radius, nsamples[level])
            assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]), f"ball query level {level}"
        x, y = ops.nearest_query(xyz, sub, m, subm), ref.nearest_query(xyz, sub, m, subm)
        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]), f"nearest level {level}"
        xyz, m, dl, radius = sub, subm, dl * 2, radius * 2


def test_gather_and_gradient_against_reference_cuda(cuda_device, ref):
    from deep3dpointclouddenoising_b200 import ops
    pts, mask, _, _ = synthetic.make_batch(950, 4, 2048, ragged=True)
    xyz, m = torch.from_numpy(pts).to(cuda_device), torch.from_numpy(mask).to(cuda_device)
    idx, _ = ops.ball_query(xyz, xyz, m, m, 0.025, 52)
    f = torch.randn(4, 24, 2048, device=cuda_device)
    assert torch.equal(ops.group_points(f, idx), ref.group_points(f, idx))
    g = torch.randn(4, 24, 2048, 52, device=cuda_device)
    torch.testing.assert_close(ops.group_points_grad(g, idx, 2048), ref.group_points_grad(g, idx, 2048), rtol=1e-4, atol=1e-3)
