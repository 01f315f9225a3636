"""GPU parity of the staged-tile PosPool kernels (csrc/pospool_tiles.cu: cp.async.bulk staging + tcgen05 contraction)
and of the Morton processing order they use (csrc/spatial_order.cu).

Checked against the float oracle (oracle/aggregation_ref.py, the reference's eager formula
u_net_arch/models/local_aggregation_operators.py:140-183 and its autograd backward) with the SAME fp32 tolerances
as the per-query gather kernels (outputs rtol 1e-5 / atol 2e-6, gradients rtol 1e-4 / atol 2e-5), at small ragged
shapes, at every channel width of the U-Net and at the benched level-0 shape (8192 points, 72 channels, 52 slots).
"""
import numpy as np
import pytest
import torch

from deep3dpointclouddenoising_b200 import synthetic
from oracle import aggregation_ref as agg

pytestmark = pytest.mark.gpu
FWD = dict(rtol=1e-5, atol=2e-6)
BWD = dict(rtol=1e-4, atol=2e-5)


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def morton_order_numpy(p):
    """Restatement of spatial_order.cu: 64^3 cells over the bounding box, key = morton << 14 | index."""
    lo = p.min(0)
    ext = np.float32((p.max(0) - lo).max())
    s = np.float32(63.999) / ext if ext > 0 else np.float32(0)
    cell = np.clip(((p - lo) * s).astype(np.int32), 0, 63).astype(np.uint32)
    code = np.zeros(len(p), np.uint32)
    for bit in range(6):
        for d in range(3):
            code |= ((cell[:, d] >> np.uint32(bit)) & np.uint32(1)) << np.uint32(3 * bit + d)
    return np.argsort((code.astype(np.uint64) << np.uint64(14)) | np.arange(len(p), dtype=np.uint64), kind="stable")


@pytest.mark.parametrize("N", [64, 300, 2048, 8192])
def test_spatial_order_is_the_morton_permutation(cuda_device, N):
    from deep3dpointclouddenoising_b200 import ops
    pts, _, _, _ = synthetic.make_batch(77 + N, 3, N, ragged=True)
    order = ops.spatial_order(dev(pts, cuda_device)).cpu().numpy()
    for b in range(pts.shape[0]):
        assert np.array_equal(np.sort(order[b]), np.arange(N))
        assert np.array_equal(order[b], morton_order_numpy(pts[b]))


def test_ball_query_winners_by_support_index(cuda_device):
    """idx_by_support: the same winners as idx[:, :, :nvalid], ascending support index, distance rank in bits 16..23."""
    from deep3dpointclouddenoising_b200 import ops
    for (N, M, ns, radius) in ((2048, 2048, 52, 0.025), (8192, 8192, 52, 0.025), (1000, 1000, 7, 0.004), (600, 600, 64, 0.2)):
        pts, mask, _, _ = synthetic.make_batch(5 + N, 2, N, ragged=True)
        d, dm = dev(pts, cuda_device), dev(mask, cuda_device)
        idx, msk, nv, bys = [t.cpu().numpy() for t in ops.ball_query(d, d, dm, dm, radius, ns, want_nvalid=True, want_by_support=True)]
        for b in range(2):
            for j in range(0, M, 37):
                n = int(nv[b, j])
                row = bys[b, j]
                assert (row[n:] == -1).all() and (row[:n] >= 0).all()
                ids, ranks = row[:n] & 0xffff, row[:n] >> 16
                assert (np.diff(ids) > 0).all()
                assert np.array_equal(idx[b, j, ranks], ids) and np.array_equal(np.sort(ranks), np.arange(n))


def _case(oracle, seed, B, N, M, ns, radius, C):
    pts, mask, _, _ = synthetic.make_batch(seed, B, N, ragged=True)
    if M == N:
        q, qm = pts, mask
    else:
        q, qm = oracle.grid_subsampling(pts, mask, M, 0.05 / 32 * (N / M) ** 0.5)
    idx, msk = oracle.ball_query(q, pts, qm, mask, radius, ns)
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((B, C, N)).astype(np.float32)
    gout = rng.standard_normal((B, C, M)).astype(np.float32)
    return pts, mask, q, qm, idx, msk, f, gout


@pytest.mark.parametrize("C,N,M,ns,radius,reduction", [
    (72, 2048, 2048, 52, 0.025, 'avg'), (72, 2048, 2048, 52, 0.025, 'sum'), (144, 2048, 512, 39, 0.03, 'avg'),
    (288, 512, 512, 32, 0.05, 'avg'), (576, 256, 64, 26, 0.08, 'avg'), (1152, 64, 64, 26, 0.4, 'avg'),
    (12, 300, 300, 7, 0.02, 'avg'), (24, 1000, 130, 20, 0.03, 'sum'), (96, 700, 700, 64, 0.05, 'avg')])
def test_staged_pospool_against_float_oracle(cuda_device, oracle, C, N, M, ns, radius, reduction):
    from deep3dpointclouddenoising_b200 import ops
    B = 2
    pts, mask, q, qm, idx, msk, f, gout = _case(oracle, 900 + C + ns, B, N, M, ns, radius, C)
    tf = torch.from_numpy(f).requires_grad_(True)
    targs = [torch.from_numpy(np.ascontiguousarray(a)) for a in (q, pts, qm, idx, msk)]
    ref = agg.pospool(tf, *targs, radius, reduction)
    (ref_g,) = torch.autograd.grad(ref, tf, torch.from_numpy(gout))
    dq, ds, dqm, dsm = dev(q, cuda_device), dev(pts, cuda_device), dev(qm, cuda_device), dev(mask, cuda_device)
    didx, dmsk, dnv, dbys = ops.ball_query(dq, ds, dqm, dsm, radius, ns, want_nvalid=True, want_by_support=True)
    assert np.array_equal(didx.cpu().numpy(), idx)
    f_cl = dev(f.transpose(0, 2, 1), cuda_device)
    g_cl = dev(gout.transpose(0, 2, 1), cuda_device)
    oq = ops.spatial_order(dq)
    out = ops.pospool_fwd(f_cl, dq, ds, didx, dnv, dqm, radius, reduction, query_order=oq, idx_by_support=dbys)
    # 'sum' is 'avg' times the neighbourhood size: the same relative accuracy means an absolute tolerance nsample times larger
    FWD = dict(rtol=1e-5, atol=2e-6 * (ns if reduction == 'sum' else 1))
    BWD = dict(rtol=1e-4, atol=2e-5 * (ns if reduction == 'sum' else 1))
    np.testing.assert_allclose(out.cpu().numpy().transpose(0, 2, 1), ref.detach().numpy(), **FWD)
    # backward, scatter form: the forward tile's transposed contraction, partial sums added with float atomics
    plan = ops.tile_plan(dbys, dnv, dqm, oq, N)
    gs = ops.pospool_bwd(g_cl, dq, ds, None, None, dnv, dqm, N, ns, radius, reduction, query_order=oq, idx_by_support=dbys,
                         plan=plan, ordered=False)
    np.testing.assert_allclose(gs.cpu().numpy().transpose(0, 2, 1), ref_g.numpy(), **BWD)
    # ordered form: the tiles' partial rows added per support in ascending tile order — no float atomics, same bits twice
    go = ops.pospool_bwd(g_cl, dq, ds, None, None, dnv, dqm, N, ns, radius, reduction, query_order=oq, idx_by_support=dbys,
                         plan=plan, ordered=True)
    np.testing.assert_allclose(go.cpu().numpy().transpose(0, 2, 1), ref_g.numpy(), **BWD)
    assert torch.equal(go, ops.pospool_bwd(g_cl, dq, ds, None, None, dnv, dqm, N, ns, radius, reduction, query_order=oq,
                                           idx_by_support=dbys, plan=plan, ordered=True))
    # the forward has a fixed summation order: same bits on a second run (also with the plan built inside the call)
    assert torch.equal(out, ops.pospool_fwd(f_cl, dq, ds, didx, dnv, dqm, radius, reduction, query_order=oq, idx_by_support=dbys,
                                            plan=plan))
    # and both agree with the per-query gather kernels
    rowptr, entries = ops.build_inverse_map(didx, N)
    legacy = ops.pospool_fwd(f_cl, dq, ds, didx, dnv, dqm, radius, reduction)
    np.testing.assert_allclose(out.cpu().numpy(), legacy.cpu().numpy(), **FWD)
    legacy_g = ops.pospool_bwd(g_cl, dq, ds, rowptr, entries, dnv, dqm, N, ns, radius, reduction)
    np.testing.assert_allclose(gs.cpu().numpy(), legacy_g.cpu().numpy(), **BWD)


def test_staged_pospool_at_the_benched_level0_shape(cuda_device, oracle):
    """BASELINE config 2, first level: 8192 points, 72 channels, 52 slots, radius 0.025 (2 of the 16 clouds)."""
    from deep3dpointclouddenoising_b200 import ops
    B, N, C, ns, radius = 2, 8192, 72, 52, 0.025
    pts, mask, q, qm, idx, msk, f, gout = _case(oracle, 31, B, N, N, ns, radius, C)
    tf = torch.from_numpy(f).requires_grad_(True)
    targs = [torch.from_numpy(np.ascontiguousarray(a)) for a in (q, pts, qm, idx, msk)]
    ref = agg.pospool(tf, *targs, radius, 'avg')
    (ref_g,) = torch.autograd.grad(ref, tf, torch.from_numpy(gout))
    ds, dsm = dev(pts, cuda_device), dev(mask, cuda_device)
    didx, dmsk, dnv, dbys = ops.ball_query(ds, ds, dsm, dsm, radius, ns, want_nvalid=True, want_by_support=True)
    assert np.array_equal(didx.cpu().numpy(), idx)
    order = ops.spatial_order(ds)
    f_cl, g_cl = dev(f.transpose(0, 2, 1), cuda_device), dev(gout.transpose(0, 2, 1), cuda_device)
    out = ops.pospool_fwd(f_cl, ds, ds, didx, dnv, dsm, radius, 'avg', query_order=order, idx_by_support=dbys)
    np.testing.assert_allclose(out.cpu().numpy().transpose(0, 2, 1), ref.detach().numpy(), **FWD)
    gs = ops.pospool_bwd(g_cl, ds, ds, None, None, dnv, dsm, N, ns, radius, 'avg', query_order=order, idx_by_support=dbys)
    np.testing.assert_allclose(gs.cpu().numpy().transpose(0, 2, 1), ref_g.numpy(), **BWD)


def test_staged_pospool_far_from_the_origin(cuda_device, oracle):
    """The tile-centred formulation must not lose accuracy when the cloud sits far from the origin."""
    from deep3dpointclouddenoising_b200 import ops
    B, N, C, ns, radius = 2, 1024, 24, 20, 0.03
    pts, mask, q, qm, idx, msk, f, gout = _case(oracle, 57, B, N, N, ns, radius, C)
    pts = (pts + np.array([3.0, -2.0, 5.0], np.float32)).astype(np.float32)
    idx, msk = oracle.ball_query(pts, pts, mask, mask, radius, ns)
    tf = torch.from_numpy(f).requires_grad_(True)
    targs = [torch.from_numpy(np.ascontiguousarray(a)) for a in (pts, pts, mask, idx, msk)]
    ref = agg.pospool(tf, *targs, radius, 'avg')
    ds, dsm = dev(pts, cuda_device), dev(mask, cuda_device)
    didx, dmsk, dnv, dbys = ops.ball_query(ds, ds, dsm, dsm, radius, ns, want_nvalid=True, want_by_support=True)
    out = ops.pospool_fwd(dev(f.transpose(0, 2, 1), cuda_device), ds, ds, didx, dnv, dsm, radius, 'avg',
                          query_order=ops.spatial_order(ds), idx_by_support=dbys)
    # coordinates ~5 carry an absolute rounding of 5e-7, i.e. ~2e-5 of radius: the reference's own fp32 result has it too
    np.testing.assert_allclose(out.cpu().numpy().transpose(0, 2, 1), ref.detach().numpy(), rtol=1e-5, atol=1e-5)
