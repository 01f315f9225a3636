"""GPU parity, index ops: the sm_100a kernels, called through the C ABI, must be BIT-EXACT with the oracle
(and with the committed outputs of the reference itself) — ordering, 3*nsample truncation, nearest swap,
cyclic padding, LCG shuffle included.  At the full BASELINE size the check is against the oracle on a
subset of clouds plus size-independent properties."""
import os

import numpy as np
import pytest
import torch

from deep3dpointclouddenoising_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ops():
    from deep3dpointclouddenoising_b200 import ops
    return ops


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def test_golden_vectors_from_the_reference(cuda_device):
    ops, g = _ops(), np.load(os.path.join(GOLD, "index_ops.npz"))
    pts, mask = dev(g["points"], cuda_device), dev(g["mask"], cuda_device)
    for name, radius, ns in (("self_r025_ns52", 0.025, 52), ("self_r05_ns16", 0.05, 16), ("self_r005_ns8", 0.005, 8)):
        idx, msk = ops.ball_query(pts, pts, mask, mask, radius, ns)
        assert np.array_equal(idx.cpu().numpy(), g[f"bq_{name}_idx"]), name
        assert np.array_equal(msk.cpu().numpy(), g[f"bq_{name}_mask"]), name
    sub, subm = ops.grid_subsample(pts, mask, 256, 0.003125)
    assert np.array_equal(sub.cpu().numpy(), g["gs_dl003125_m256_xyz"]) and np.array_equal(subm.cpu().numpy(), g["gs_dl003125_m256_mask"])
    sub2, subm2 = ops.grid_subsample(pts, mask, 1024, 0.0125)
    assert np.array_equal(sub2.cpu().numpy(), g["gs_dl0125_m1024_xyz"]) and np.array_equal(subm2.cpu().numpy(), g["gs_dl0125_m1024_mask"])
    idx, msk = ops.ball_query(sub, pts, subm, mask, 0.025, 52)
    assert np.array_equal(idx.cpu().numpy(), g["bq_sub_r025_ns52_idx"]) and np.array_equal(msk.cpu().numpy(), g["bq_sub_r025_ns52_mask"])
    nidx, nmsk = ops.nearest_query(pts, sub, mask, subm)
    assert np.array_equal(nidx.cpu().numpy(), g["nn_idx"]) and np.array_equal(nmsk.cpu().numpy(), g["nn_mask"])


@pytest.mark.parametrize("seed,n,ragged", [(0, 37, True), (1, 64, False), (2, 500, True), (3, 1000, True), (4, 2048, True),
                                           (5, 3001, True)])
def test_pyramid_against_oracle(cuda_device, oracle, seed, n, ragged):
    """grid subsample -> strided ball query -> self ball query -> nearest, all levels, ragged masks."""
    ops = _ops()
    pts, mask, _, _ = synthetic.make_batch(200 + seed, 3, n, ragged=ragged)
    if seed == 2:
        pts[0, 10:20] = pts[0, 9]  # duplicates: zero-distance ties
    if seed == 3:
        mask[1, :] = 0  # empty cloud: defined as idx 0 / mask 0 (reference: undefined, i % 0)
    dl, radius = 0.003125, 0.025
    xyz, m = pts, mask
    for level, ns in enumerate((52, 39, 32)):
        npoint = max(xyz.shape[1] // 4, 1)
        dx, dm = dev(xyz, cuda_device), dev(m, cuda_device)
        sub, subm = ops.grid_subsample(dx, dm, npoint, dl)
        o_sub, o_subm = oracle.grid_subsampling(xyz, m, npoint, dl)
        assert np.array_equal(sub.cpu().numpy(), o_sub), f"level {level} sub_xyz"
        assert np.array_equal(subm.cpu().numpy(), o_subm), f"level {level} sub_mask"
        for q, qm, dq, dqm in ((xyz, m, dx, dm), (o_sub, o_subm, sub, subm)):
            idx, msk, nv = ops.ball_query(dq, dx, dqm, dm, radius, ns, want_nvalid=True)
            o_idx, o_msk = oracle.ball_query(q, xyz, qm, m, radius, ns)
            assert np.array_equal(idx.cpu().numpy(), o_idx), f"level {level} idx"
            assert np.array_equal(msk.cpu().numpy(), o_msk), f"level {level} idx_mask"
            valid_q = qm.astype(bool)
            assert np.array_equal(nv.cpu().numpy()[valid_q], o_msk.sum(-1)[valid_q])
        nidx, nmsk = ops.nearest_query(dx, sub, dm, subm)
        o_nidx, o_nmsk = oracle.nearest_query(xyz, o_sub, m, o_subm)
        assert np.array_equal(nidx.cpu().numpy(), o_nidx) and np.array_equal(nmsk.cpu().numpy(), o_nmsk)
        xyz, m, dl, radius = o_sub, o_subm, dl * 2, radius * 2


@pytest.mark.parametrize("ns", [1, 3, 33, 64, 100])
def test_ball_query_nsample_range(cuda_device, oracle, ns):
    ops = _ops()
    pts, mask, _, _ = synthetic.make_batch(300 + ns, 2, 700, ragged=True)
    dx, dm = dev(pts, cuda_device), dev(mask, cuda_device)
    for radius in (0.004, 0.02, 0.2):  # sparse (cyclic padding) ... everything in radius (3*ns truncation + swap)
        idx, msk = ops.ball_query(dx, dx, dm, dm, radius, ns)
        o_idx, o_msk = oracle.ball_query(pts, pts, mask, mask, radius, ns)
        assert np.array_equal(idx.cpu().numpy(), o_idx) and np.array_equal(msk.cpu().numpy(), o_msk)


def test_full_size_level0_against_oracle_subset_and_properties(cuda_device, oracle):
    """BASELINE size: B=16 x 8192, r=0.025, ns=52.  Oracle on 2 clouds; properties on all 16."""
    ops = _ops()
    B, N, ns, radius = 16, 8192, 52, 0.025
    pts, mask, _, _ = synthetic.make_batch(1234, B, N, ragged=True)
    dx, dm = dev(pts, cuda_device), dev(mask, cuda_device)
    idx, msk, nv = ops.ball_query(dx, dx, dm, dm, radius, ns, want_nvalid=True)
    idx_h, msk_h = idx.cpu().numpy(), msk.cpu().numpy()
    for b in (0, 9):
        o_idx, o_msk = oracle.ball_query(pts[b:b + 1], pts[b:b + 1], mask[b:b + 1], mask[b:b + 1], radius, ns)
        assert np.array_equal(idx_h[b], o_idx[0]) and np.array_equal(msk_h[b], o_msk[0])
    # properties: indices among valid supports; slot 0 of a self query is a zero-distance point; distances ascend
    v = mask.sum(1)
    assert (idx_h < v[:, None, None]).all() and (idx_h >= 0).all()
    for b in range(B):
        nb = pts[b][idx_h[b]]  # (N, ns, 3)
        dist = ((nb - pts[b][:, None, :]) ** 2).sum(-1)
        assert (dist[:, 0] == 0).all()
        valid = msk_h[b].astype(bool)
        asc = (np.diff(dist, axis=1) >= -1e-9) | ~valid[:, 1:]  # numpy recomputes d2 without the fma: allow its rounding
        assert asc.all()
        assert (dist[valid] < radius * radius * (1 + 1e-5)).all()
    # determinism: a second launch gives the same bits
    idx2, msk2 = ops.ball_query(dx, dx, dm, dm, radius, ns)
    assert torch.equal(idx, idx2) and torch.equal(msk, msk2)
    sub, subm = ops.grid_subsample(dx, dm, 2048, 0.003125)
    o_sub, o_subm = oracle.grid_subsampling(pts, mask, 2048, 0.003125)
    assert np.array_equal(sub.cpu().numpy(), o_sub) and np.array_equal(subm.cpu().numpy(), o_subm)


def test_large_cloud_uses_global_sort_buffer(cuda_device, oracle):
    ops = _ops()
    pts = synthetic.make_cloud(0, 20000)[None] * 0.1
    mask = np.ones((1, 20000), np.int32)
    sub, subm = ops.grid_subsample(dev(pts, cuda_device), dev(mask, cuda_device), 6000, 0.002)
    o_sub, o_subm = oracle.grid_subsampling(pts, mask, 6000, 0.002)
    assert np.array_equal(sub.cpu().numpy(), o_sub) and np.array_equal(subm.cpu().numpy(), o_subm)


def test_group_points_and_grad(cuda_device, oracle):
    ops = _ops()
    rng = np.random.default_rng(0)
    for (B, C, N, M, ns) in ((2, 7, 300, 50, 9), (3, 24, 1000, 1000, 16), (1, 3, 64, 64, 5)):
        pts, mask, _, _ = synthetic.make_batch(400 + N, B, N, ragged=True)
        q = pts[:, :M]
        idx, _ = oracle.ball_query(q, pts, mask[:, :M], mask, 0.03, ns)
        f = rng.standard_normal((B, C, N)).astype(np.float32)
        out = ops.group_points(dev(f, cuda_device), dev(idx, cuda_device))
        assert np.array_equal(out.cpu().numpy(), oracle.group_points(f, idx))  # a copy: exact
        g = rng.standard_normal((B, C, M, ns)).astype(np.float32)
        dg, di = dev(g, cuda_device), dev(idx, cuda_device)
        gp = ops.group_points_grad(dg, di, N, deterministic=True)
        # tolerance: fp32 sum of <= a few hundred terms in a different order than the oracle's exact sum
        np.testing.assert_allclose(gp.cpu().numpy(), oracle.group_points_grad(g, idx, N), rtol=1e-4, atol=1e-4)
        assert torch.equal(gp, ops.group_points_grad(dg, di, N, deterministic=True))  # fixed order, no float atomics
        fast = ops.group_points_grad(dg, di, N)  # default: shared-memory atomics per (b, c) plane, like the reference
        np.testing.assert_allclose(fast.cpu().numpy(), oracle.group_points_grad(g, idx, N), rtol=1e-4, atol=1e-4)


def test_inverse_map_is_sorted_and_complete(cuda_device, oracle):
    ops = _ops()
    pts, mask, _, _ = synthetic.make_batch(77, 4, 2048, ragged=True)
    idx, _ = oracle.ball_query(pts, pts, mask, mask, 0.025, 52)
    rowptr, entries = ops.build_inverse_map(dev(idx, cuda_device), 2048)
    rp, en = rowptr.cpu().numpy(), entries.cpu().numpy()
    B, N, ns = 4, 2048, 52
    assert rp[0] == 0 and rp[-1] == B * N * ns and (np.diff(rp) >= 0).all()
    counts = np.stack([np.bincount(idx[b].ravel(), minlength=N) for b in range(B)]).ravel()
    assert np.array_equal(np.diff(rp), counts)
    for row in list(range(0, 64)) + [N + 3, 3 * N + 17]:
        seg = en[rp[row]:rp[row + 1]]
        b, i = divmod(row, N)
        assert (np.diff(seg) > 0).all()  # ascending (query, slot), unique
        assert (idx[b][seg >> 8, seg & 255] == i).all()
