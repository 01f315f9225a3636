"""GPU parity, row f2: patch centres, radius patches and vote averaging against the reference's own building blocks
(its CPU grid_subsampling.cpp through oracle/_ref/libref_gridsub.so when present, sklearn's KDTree — the class the
reference dataset uses — and numpy vote accumulation as in qualitative_inference_test.py:325-342)."""
import os

import numpy as np
import pytest
import torch
from sklearn.neighbors import KDTree

from deep3dpointclouddenoising_b200 import synthetic

pytestmark = pytest.mark.gpu


def _cloud(n, seed):
    return synthetic.make_cloud(seed, n, sigma=0.004)


def _numpy_barycentres(pts, dl):
    """grid_subsampling.cpp:25-103 restated (float32), voxels in ascending id."""
    dl32 = np.float32(dl)
    inv = np.float32(1.0) / dl32
    origin = (np.floor(pts.min(0) * inv) * dl32).astype(np.float32)
    mx = pts.max(0)
    nx = int(np.floor((mx[0] - origin[0]) / dl32)) + 1
    ny = int(np.floor((mx[1] - origin[1]) / dl32)) + 1
    ijk = np.floor((pts - origin) / dl32).astype(np.int64)
    ids = ijk[:, 0] + nx * ijk[:, 1] + nx * ny * ijk[:, 2]
    out = []
    for v in np.unique(ids):
        s = np.zeros(3, np.float32)
        members = np.nonzero(ids == v)[0]
        for i in members:  # ascending point index, fp32 accumulation like SampledData::update_points
            s = (s + pts[i]).astype(np.float32)
        out.append(s * np.float32(1.0 / len(members)))
    return np.stack(out)


def test_voxel_barycentres_match_reference_cpu(cuda_device):
    from deep3dpointclouddenoising_b200 import inference
    from oracle import cpu_index_ops
    pts = _cloud(6000, 1)
    bary, counts = inference.voxel_barycentres(torch.from_numpy(pts).to(cuda_device), 0.05)
    got = bary.cpu().numpy()
    assert np.array_equal(got, _numpy_barycentres(pts, 0.05))
    assert int(counts.sum()) == 6000
    try:
        ref = cpu_index_ops.ref_gridsub_cpu().compute(pts, 0.05)  # the reference's own C++ (hash-map order)
    except (FileNotFoundError, OSError):
        return
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(key(got), key(ref))


def test_config0_pyramid_on_100k_cloud(cuda_device):
    """BASELINE configs[0] at the geometry BASELINE.md §3 states (tools/pyramid_100k.py): subsampling at
    dl = 0.0015625 * 2^l, the U-Net's nine radius lists (r = 0.025 * 2^(l-1) strided, 0.025 * 2^l self, nsample nearest)
    of one 100k-point noisy shape.  Every level's barycentres equal the reference's grid_subsampling.cpp bit for bit (as
    sets: the reference emits hash-map order); the radius lists equal the reference's vendored nanoflann radiusSearch
    (oracle/_ref/libref_nanoflann.so) or, where that did not travel, sklearn's KDTree — the reference's host search."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import pyramid_100k as P
    from deep3dpointclouddenoising_b200 import inference
    from oracle import cpu_index_ops
    pts = _cloud(100_000, 0)
    try:
        ref = cpu_index_ops.ref_gridsub_cpu()
    except (FileNotFoundError, OSError):
        ref = None
    try:
        nf = cpu_index_ops.ref_nanoflann()
    except (FileNotFoundError, OSError):
        nf = None
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    levels_np, levels_gpu = [pts], [torch.from_numpy(pts).to(cuda_device)]
    for l in range(1, 5):
        dl = P.BASE_DL * 2 ** l
        sub, counts = inference.voxel_barycentres(levels_gpu[-1], dl)
        got = sub.cpu().numpy()
        assert int(counts.sum()) == levels_gpu[-1].shape[0]
        if ref is not None:
            want = ref.compute(levels_np[-1], dl)
            assert np.array_equal(key(got), key(want)), l
            got = want  # the next level sums points in input order (fp32): both sides continue from the SAME array
        levels_np.append(got)
        levels_gpu.append(torch.from_numpy(got).to(cuda_device))
    sizes = [len(a) for a in levels_np]
    assert sizes == sorted(sizes, reverse=True) and sizes[-1] >= 1
    for lq, ls, radius, cap in P.list_specs():
        q, s_ = levels_np[lq], levels_np[ls]
        probe = np.arange(0, len(q), max(len(q) // 200, 1))
        idx, cnt = P.radius_lists(levels_gpu[ls], levels_gpu[lq][torch.from_numpy(probe).to(cuda_device)].contiguous(), radius, cap)
        idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
        if nf is not None:
            ref_idx, ref_cnt = nf.radius(s_, q[probe], radius, cap, 4)
        else:
            ri, _ = KDTree(s_).query_radius(q[probe], r=radius, return_distance=True, sort_results=True)
            ref_cnt = np.array([len(r) for r in ri])
            ref_idx = np.stack([np.pad(r[:cap], (0, max(cap - len(r), 0)), constant_values=-1) for r in ri])
        assert np.abs(cnt - ref_cnt).max() <= 1, (lq, ls)  # a support exactly on the sphere may round either way
        for k in range(len(probe)):
            take = min(int(cnt[k]), int(ref_cnt[k]), cap)
            if not np.array_equal(idx[k, :take], ref_idx[k, :take]):  # equal-distance groups may be ordered differently
                d_got = np.linalg.norm(s_[idx[k, :take]].astype(np.float64) - q[probe[k]], axis=1)
                d_ref = np.linalg.norm(s_[ref_idx[k, :take]].astype(np.float64) - q[probe[k]], axis=1)
                np.testing.assert_allclose(d_got, d_ref, rtol=1e-5, atol=1e-9)


def test_radius_patches_match_kdtree(cuda_device):
    from deep3dpointclouddenoising_b200 import inference, ops
    pts = _cloud(30000, 2)
    dpts = torch.from_numpy(pts).to(cuda_device)
    centre_idx = inference.patch_centres(dpts, 0.05)
    tree = KDTree(pts)
    # centres: the nearest real point to every barycentre (offset_dataset.py:553 tree.query(sub_pc, k=1))
    bary = inference.voxel_barycentres(dpts, 0.05)[0].cpu().numpy()
    ref_centres = tree.query(bary, k=1, return_distance=False)[:, 0]
    mine = centre_idx.cpu().numpy()
    differ = np.nonzero(mine != ref_centres)[0]  # only exact distance ties (duplicate points) may resolve differently:
    assert len(differ) < 0.01 * len(mine)        # ours -> lowest index, sklearn -> unspecified
    d_mine = np.linalg.norm(pts[mine[differ]].astype(np.float64) - bary[differ], axis=1)
    d_ref = np.linalg.norm(pts[ref_centres[differ]].astype(np.float64) - bary[differ], axis=1)
    assert np.array_equal(d_mine, d_ref) and (mine[differ] < ref_centres[differ]).all()
    for num_points in (512, 4096):
        idx, cnt = ops.radius_patches(dpts, dpts[centre_idx].contiguous(), 0.05, num_points)
        idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
        ref_idx, ref_d = tree.query_radius(pts[centre_idx.cpu().numpy()], r=0.05, return_distance=True, sort_results=True)
        for p in range(0, len(ref_idx), 7):
            assert cnt[p] == len(ref_idx[p])
            take = min(cnt[p], num_points)
            got = idx[p, :take]
            if not np.array_equal(got, ref_idx[p][:take]):  # only the order inside groups of equal distance may differ
                d_got = np.linalg.norm(pts[got].astype(np.float64) - pts[centre_idx[p].item()], axis=1)
                np.testing.assert_allclose(d_got, ref_d[p][:take], rtol=1e-12, atol=1e-12)
            assert (idx[p, take:] == -1).all()


def test_extract_and_vote(cuda_device):
    from deep3dpointclouddenoising_b200 import inference, ops
    pts = _cloud(20000, 3)
    dpts = torch.from_numpy(pts).to(cuda_device)
    centre_idx = inference.patch_centres(dpts, 0.05)
    n_pts = 1024
    p, mask, feats, inds = inference.extract_patches(dpts, centre_idx, 0.05, n_pts, seed=1)
    P = p.shape[0]
    assert torch.equal(inds[:, 0], centre_idx) and (p[:, 0].abs().max() == 0)
    assert torch.equal(feats, p.transpose(1, 2)) and bool((mask.sum(1) >= 1).all())
    # every valid slot is a distinct in-radius point, every padded slot repeats a valid one
    d = (dpts[inds] - dpts[centre_idx][:, None]).norm(dim=2)
    assert bool((d <= 0.05 + 1e-6).all())
    for r in range(0, P, 11):
        v = inds[r][mask[r].bool()]
        assert v.unique().numel() == v.numel()
        assert set(inds[r][~mask[r].bool()].tolist()) <= set(v.tolist())
    # vote averaging with a fake "prediction" = patch id in x, slot in y, 1 in z
    pred = torch.zeros(P, 3, n_pts, device=cuda_device)
    pred[:, 0] = torch.arange(P, device=cuda_device)[:, None].float()
    pred[:, 1] = torch.arange(n_pts, device=cuda_device)[None, :].float()
    pred[:, 2] = 1.0
    flat = torch.arange(P * n_pts, device=cuda_device).view(P, n_pts)
    pool = 1 << 20
    vote_idx = torch.where(mask.bool(), inds, 20000 + (flat & (pool - 1))).int().view(1, P * n_pts // 128, 128).contiguous()
    rowptr, entries = ops.build_inverse_map(vote_idx, 20000 + pool)
    mean, votes = ops.vote_mean(pred, rowptr, entries, 20000, n_pts)
    s = np.zeros((20000, 3), np.float32)
    c = np.zeros((20000, 1), np.float32) + 1e-7
    inds_h, mask_h, pred_h = inds.cpu().numpy(), mask.cpu().numpy().astype(bool), pred.cpu().numpy()
    for b in range(P):  # qualitative_inference_test.py:325-337
        s[inds_h[b][mask_h[b]]] += pred_h[b][:, mask_h[b]].T
        c[inds_h[b][mask_h[b]]] += 1
    np.testing.assert_allclose(mean.cpu().numpy(), s / c, rtol=1e-5, atol=1e-5)
    assert np.array_equal(votes.cpu().numpy(), np.round(c[:, 0]))


def test_denoise_cloud_graph_replay_equals_eager(cuda_device):
    """Full-shape inference (row f2): the CUDA-graph replay of the per-batch forward gives the eager result."""
    import bench
    from deep3dpointclouddenoising_b200 import inference
    pts = torch.from_numpy(_cloud(60000, 5)).to(cuda_device)
    model, _, _ = bench.build_model("pospool", 1024)
    model = model.to(cuda_device).eval()
    out_g = inference.denoise_cloud(model, pts, 0.05, 0.05, 1024, batch_size=4, use_graph=True)
    out_e = inference.denoise_cloud(model, pts, 0.05, 0.05, 1024, batch_size=4, use_graph=False)
    assert torch.equal(out_g[2], out_e[2]) and bool((out_g[2] > 0).all())  # votes: every point covered
    torch.testing.assert_close(out_g[1], out_e[1], rtol=1e-5, atol=1e-6)


def test_radius_lists_histogram_selection_equals_full_sort(cuda_device):
    """Few results out of many candidates take the histogram-selection path of the radius kernel; asking for many results
    takes the full sort — the first entries must be identical (same distance order, ties by index), in every storage tier."""
    from deep3dpointclouddenoising_b200 import ops
    pts = torch.from_numpy(_cloud(40000, 11)).to(cuda_device)
    centres = pts[::97].contiguous()
    full, cnt_full = ops.radius_patches(pts, centres, 0.12, 4096)              # n <= 4 * 4096: sorted as a whole
    assert int(cnt_full.max()) > 600 and int(cnt_full.max()) <= 4096
    for kw in (dict(smem_keys=12288), dict(smem_keys=4096), dict(overflow_stride=8192)):
        few, cnt = ops.radius_patches(pts, centres, 0.12, 32, **kw)           # 32 of ~1000: selected, then sorted
        assert torch.equal(cnt, cnt_full)
        assert torch.equal(few, full[:, :32])
    auto, cnt = ops.radius_neighbors(pts, centres, 0.12, 32)                   # tier chosen from a probe
    assert torch.equal(auto, full[:, :32]) and torch.equal(cnt, cnt_full)
