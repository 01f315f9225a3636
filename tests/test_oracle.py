"""CPU suite, part 1: the oracle is pinned against outputs of the reference itself.

  * tests/golden/index_ops.npz   — produced by the reference's CUDA kernel definitions compiled for the host
  * tests/golden/aggregation.npz — produced by the reference's own Python modules (see oracle/make_golden.py)
  * when oracle/_ref/libref_emul.so exists (build container), randomised cross-checks on top.
The reference ships no tests or golden vectors of its own for this path.
"""
import os

import numpy as np
import pytest
import torch

from deep3dpointclouddenoising_b200 import synthetic
from oracle import aggregation_ref as agg
from oracle import cpu_index_ops

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold_idx():
    return np.load(os.path.join(GOLD, "index_ops.npz"))


@pytest.fixture(scope="module")
def gold_agg():
    return np.load(os.path.join(GOLD, "aggregation.npz"))


@pytest.mark.parametrize("name,radius,ns", [("self_r025_ns52", 0.025, 52), ("self_r05_ns16", 0.05, 16),
                                            ("self_r005_ns8", 0.005, 8)])
def test_ball_query_matches_reference_golden(oracle, gold_idx, name, radius, ns):
    pts, mask = gold_idx["points"], gold_idx["mask"]
    idx, msk = oracle.ball_query(pts, pts, mask, mask, radius, ns)
    assert np.array_equal(idx, gold_idx[f"bq_{name}_idx"])
    assert np.array_equal(msk, gold_idx[f"bq_{name}_mask"])


def test_grid_subsampling_nearest_and_strided_query_match_reference_golden(oracle, gold_idx):
    pts, mask = gold_idx["points"], gold_idx["mask"]
    sub, subm = oracle.grid_subsampling(pts, mask, 256, 0.003125)
    assert np.array_equal(sub, gold_idx["gs_dl003125_m256_xyz"]) and np.array_equal(subm, gold_idx["gs_dl003125_m256_mask"])
    sub2, subm2 = oracle.grid_subsampling(pts, mask, 1024, 0.0125)
    assert np.array_equal(sub2, gold_idx["gs_dl0125_m1024_xyz"]) and np.array_equal(subm2, gold_idx["gs_dl0125_m1024_mask"])
    assert subm2.sum() < subm2.size  # the padding branch is exercised
    idx, msk = oracle.ball_query(sub, pts, subm, mask, 0.025, 52)
    assert np.array_equal(idx, gold_idx["bq_sub_r025_ns52_idx"]) and np.array_equal(msk, gold_idx["bq_sub_r025_ns52_mask"])
    nidx, nmsk = oracle.nearest_query(pts, sub, mask, subm)
    assert np.array_equal(nidx, gold_idx["nn_idx"]) and np.array_equal(nmsk, gold_idx["nn_mask"])


@pytest.mark.skipif(not cpu_index_ops.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(6))
def test_restatement_equals_host_compiled_reference_kernels(oracle, seed):
    ref = cpu_index_ops.reference()
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 700))
    pts, mask, _, _ = synthetic.make_batch(100 + seed, 3, n, ragged=True)
    if seed == 0:
        pts[0, 5:9] = pts[0, 4]  # duplicate points: d2 == 0 ties resolved by index
        pts[0, mask[0] == 0] = pts[0, 4]  # padding rows must stay duplicates of VALID points (else cnt == 0)
    if seed == 1:
        mask[1, :] = 0  # an empty cloud
    dl = float(rng.choice([0.003125, 0.00625, 0.0125]))
    m = int(rng.integers(8, n))
    a, b = oracle.grid_subsampling(pts, mask, m, dl), ref.grid_subsampling(pts, mask, m, dl)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    sub, subm = a
    # cnt == 0 is undefined behaviour in the reference (i % 0 -> SIGFPE on the host build; defined as
    # idx 0 / mask 0 in the oracle and the CUDA path), so radii are chosen to keep >= 1 neighbour:
    # a cell centroid lies within sqrt(3)*dl of its members, a (padded) point within 0 of itself.
    for q, qm, radii in ((pts, mask, ((0.025, 13), (0.004, 5), (0.06, 32))), (sub, subm, ((2 * dl, 13), (5 * dl, 32)))):
        for radius, ns in radii:
            if seed == 1:
                continue
            x, y = oracle.ball_query(q, pts, qm, mask, radius, ns), ref.ball_query(q, pts, qm, mask, radius, ns)
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1])
    x, y = oracle.nearest_query(pts, sub, mask, subm), ref.nearest_query(pts, sub, mask, subm)
    assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1])
    idx = oracle.ball_query(sub, pts, subm, mask, 0.03, 9)[0]
    f = rng.standard_normal((3, 7, n)).astype(np.float32)
    assert np.array_equal(oracle.group_points(f, idx), ref.group_points(f, idx))
    g = rng.standard_normal((3, 7, m, 9)).astype(np.float32)
    np.testing.assert_allclose(oracle.group_points_grad(g, idx, n), ref.group_points_grad(g, idx, n), rtol=1e-5, atol=1e-5)


def test_ball_query_edge_cases(oracle):
    # cnt < nsample -> cyclic padding with mask 0; padded query -> mask row 0 but indices kept; truncation at 3*ns
    pts = np.zeros((1, 40, 3), np.float32)
    pts[0, :, 0] = np.arange(40) * 0.01
    mask = np.ones((1, 40), np.int32)
    mask[0, 30:] = 0
    idx, msk = oracle.ball_query(pts, pts, mask, mask, 0.025, 4)
    assert idx[0, 0].tolist() == [0, 1, 2, 0] and msk[0, 0].tolist() == [1, 1, 1, 0]
    assert msk[0, 35].sum() == 0 and idx[0, 35].max() < 30  # padded query, neighbours among valid supports only
    idx, msk = oracle.ball_query(pts, pts, mask, mask, 1.0, 2)  # everything in radius: first 6 by index + nearest swap
    assert idx[0, 20].tolist() == [20, 4] and idx[0, 3].tolist() == [3, 2]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("tag", ["self", "strided"])
def test_aggregation_restatement_matches_reference_python(oracle, gold_agg, tag):
    g = gold_agg
    B, N, C, ns = [int(x) for x in g["meta"]]
    radius = float(g["radius"])
    pts, mask = g["points"], g["mask"]
    q, qm = (pts, mask) if tag == "self" else (g["sub_xyz"], g["sub_mask"])
    idx, msk = oracle.ball_query(q, pts, qm, mask, radius, ns)
    xyz, qx, m_, qm_, idx_, msk_ = _t(pts), _t(q), _t(mask), _t(qm), _t(idx), _t(msk)
    for red, emb, key in (("avg", "xyz", "avg"), ("sum", "xyz", "sum"), ("max", "xyz", "max"), ("avg", "sin_cos", "sincos")):
        f = _t(g["features"]).requires_grad_(True)
        y = agg.pospool(f, qx, xyz, qm_, idx_, msk_, radius, red, emb)
        np.testing.assert_allclose(y.detach().numpy(), g[f"pospool_{tag}_{key}_out"], rtol=1e-5, atol=1e-6)
        (gf,) = torch.autograd.grad(y, f, _t(g[f"pospool_{tag}_{key}_gout"]))
        np.testing.assert_allclose(gf.numpy(), g[f"pospool_{tag}_{key}_gfeat"], rtol=1e-4, atol=1e-5)
    for infl in ("linear", "constant"):
        pre = f"pseudogrid_{tag}_{infl}"
        f = _t(g["features"]).requires_grad_(True)
        w = _t(g[pre + "_weights"]).requires_grad_(True)
        y = agg.pseudogrid(f, w, _t(g[pre + "_kpoints"]), qx, xyz, qm_, idx_, msk_, float(g[pre + "_extent"]), infl)
        np.testing.assert_allclose(y.detach().numpy(), g[pre + "_out"], rtol=1e-5, atol=1e-6)
        gf, gw = torch.autograd.grad(y, (f, w), _t(g[pre + "_gout"]))
        np.testing.assert_allclose(gf.numpy(), g[pre + "_gfeat"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(gw.numpy(), g[pre + "_gweights"], rtol=1e-4, atol=1e-4)


def test_maxpool_and_upsample_restatement_match_reference_python(oracle, gold_agg):
    g = gold_agg
    B, N, C, ns = [int(x) for x in g["meta"]]
    radius = float(g["radius"])
    pts, mask = g["points"], g["mask"]
    sub, subm = oracle.grid_subsampling(pts, mask, 96, 0.00625)
    assert np.array_equal(sub, g["maxpool_sub_xyz"]) and np.array_equal(subm, g["maxpool_sub_mask"])
    idx, _ = oracle.ball_query(sub, pts, subm, mask, radius, ns)
    f = _t(g["features"]).requires_grad_(True)
    y = agg.max_pool(f, _t(idx))
    np.testing.assert_array_equal(y.detach().numpy(), g["maxpool_out"])
    (gf,) = torch.autograd.grad(y, f, _t(g["maxpool_gout"]))
    np.testing.assert_allclose(gf.numpy(), g["maxpool_gfeat"], rtol=1e-5, atol=1e-6)
    nidx, _ = oracle.nearest_query(pts, sub, mask, subm)
    cf = _t(g["upsample_features"]).requires_grad_(True)
    y = agg.nearest_upsample(cf, _t(nidx))
    np.testing.assert_array_equal(y.detach().numpy(), g["upsample_out"])
    (gf,) = torch.autograd.grad(y, cf, _t(g["upsample_gout"]))
    np.testing.assert_allclose(gf.numpy(), g["upsample_gfeat"], rtol=1e-5, atol=1e-6)


def test_nanoflann_harness_matches_brute_force():
    """oracle/_ref/libref_nanoflann.so (the reference's vendored nanoflann.hpp + PointCloud adaptor): nearest-first radius
    lists equal a float64 brute force (the CPU neighbour baseline of bench.py / tools/pyramid_100k.py)."""
    from oracle import cpu_index_ops
    try:
        nf = cpu_index_ops.ref_nanoflann()
    except (FileNotFoundError, OSError):
        pytest.skip("libref_nanoflann.so not built (no /root/reference at build time)")
    rng = np.random.default_rng(3)
    s = rng.uniform(-1, 1, (2000, 3)).astype(np.float32)
    q = rng.uniform(-1, 1, (150, 3)).astype(np.float32)
    for threads in (1, 4):
        idx, cnt = nf.radius(s, q, 0.3, 24, threads)
        d2 = ((q[:, None, :].astype(np.float64) - s[None].astype(np.float64)) ** 2).sum(-1)
        for j in range(len(q)):
            inside = np.nonzero(d2[j] < np.float32(0.3) ** 2)[0]
            assert abs(cnt[j] - len(inside)) <= 1  # a support exactly on the sphere may round either way in fp32
            order = inside[np.argsort(d2[j, inside], kind="stable")][:24]
            got = idx[j][idx[j] >= 0]
            assert len(got) == min(cnt[j], 24)
            np.testing.assert_allclose(d2[j, got], d2[j, order[:len(got)]], rtol=1e-5, atol=1e-9)
