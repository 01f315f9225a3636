"""Full-shape denoising on the GPU: patch centres, radius patches, batched U-Net inference,
vote averaging, Chamfer evaluation — the pipeline of the reference's qualitative_inference_test.py / offset_dataset.py
(test split) / compute_cd.py without leaving the device.

    ref: offset_dataset.py:540-561            centres = grid_subsampling(cloud, sample_Dl_patches) -> nearest real point
    ref: offset_dataset.py:630-733            patch = sorted radius query, first num_points, shuffle, pad, centre first
    ref: u_net_arch/qualitative_inference_test.py:282-344   per-point vote sum / (count + 1e-7)
    ref: u_net_arch/compute_cd.py:74-75       Chamfer(clean, denoised) / Chamfer(clean, noisy)
"""
import numpy as np
import torch

from . import ops


def _voxel_geometry(points, dl):
    """origin / NX / NY / NZ exactly as grid_subsampling.cpp:25-31 computes them (fp32 arithmetic)."""
    mn = points.min(0).values.cpu().numpy().astype(np.float32)
    mx = points.max(0).values.cpu().numpy().astype(np.float32)
    dl32 = np.float32(dl)
    inv = np.float32(1.0) / dl32
    origin = (np.floor(mn * inv) * dl32).astype(np.float32)
    dims = [int(np.floor((mx[d] - origin[d]) / dl32)) + 1 for d in range(3)]
    return origin, dims


def voxel_barycentres(points, dl):
    """Barycentre of every occupied voxel, ascending voxel id (the reference emits them in unordered_map order)."""
    n = points.shape[0]
    origin, (nx, ny, nz) = _voxel_geometry(points, dl)
    n_cells = nx * ny * nz
    ids = ops.voxel_ids(points, origin, dl, nx, ny, n_cells)
    if n_cells > 4 * n:
        # fine grids are almost empty (320^3 cells for 100k points at dl = 0.003): rank the occupied voxels instead of
        # walking every cell — same barycentres in the same (ascending voxel id) order, 149 -> 1 ms
        occupied, compact = torch.unique(ids, sorted=True, return_inverse=True)
        m = int(occupied.shape[0])
        rowptr, entries = ops.build_inverse_map(compact.to(torch.int32).view(1, n, 1), m)
        return ops.voxel_barycentres(points, rowptr, entries, m)
    rowptr, entries = ops.build_inverse_map(ids.view(1, n, 1), n_cells)
    bary, counts = ops.voxel_barycentres(points, rowptr, entries, n_cells)
    keep = counts > 0
    return bary[keep], counts[keep]


def patch_centres(points, sample_dl_patches=0.05):
    """Indices of the patch centres: the real point nearest to each voxel barycentre (offset_dataset.py:547-553)."""
    bary, _ = voxel_barycentres(points, sample_dl_patches)
    _, idx = ops.nn_sqdist(bary, points, want_idx=True, precise=True)  # fp64 decisions like the reference's KD-tree
    return idx.long()


def extract_patches(points, centre_idx, in_radius, num_points, seed=0):
    """-> points (P, n, 3) relative to the centre, mask (P, n) int32, features (P, 3, n), input_inds (P, n) int64."""
    dev = points.device
    P = centre_idx.shape[0]
    centres = points[centre_idx].contiguous()
    idx, cnt = ops.radius_patches(points, centres, in_radius, num_points)
    cmax = int(cnt.max())
    if cmax > 12288:  # candidate list did not fit in shared memory: sort in global scratch
        stride = 1 << int(np.ceil(np.log2(cmax)))
        idx, cnt = ops.radius_patches(points, centres, in_radius, num_points, overflow_stride=stride)
    n_valid = cnt.clamp(max=num_points).long()
    slot = torch.arange(num_points, device=dev)[None, :]
    valid = slot < n_valid[:, None]
    g = torch.Generator(device=dev).manual_seed(seed)
    # shuffle the valid entries (offset_dataset.py:650,655), pad with random valid ones (:658-659)
    key = torch.rand((P, num_points), generator=g, device=dev).masked_fill(~valid, 2.0)
    shuffled = torch.gather(idx.long(), 1, key.argsort(dim=1))
    choice = (torch.rand((P, num_points), generator=g, device=dev) * n_valid[:, None]).long().clamp(max=num_points - 1)
    choice = torch.minimum(choice, (n_valid[:, None] - 1).clamp(min=0))
    input_inds = torch.where(valid, shuffled, torch.gather(shuffled, 1, choice))
    # the centre goes to slot 0 (:683)
    pos = (input_inds == centre_idx[:, None]).float().argmax(dim=1)
    rows = torch.arange(P, device=dev)
    first = input_inds[:, 0].clone()
    input_inds[:, 0] = input_inds[rows, pos]
    input_inds[rows, pos] = first
    pts = (points[input_inds] - centres[:, None, :]).contiguous()
    return pts, valid.int().contiguous(), pts.transpose(1, 2).contiguous(), input_inds


class _GraphedForward:
    """One batch-sized forward of the model captured as a CUDA graph (static input buffers, replayed per batch): the
    eager forward is ~400 launches and bound by the host launch rate (8-10 ms per batch), the replay by the GPU (~2 ms)."""

    def __init__(self, model, batch_size, num_points, device):
        self.inputs = (torch.zeros((batch_size, num_points, 3), device=device),
                       torch.ones((batch_size, num_points), dtype=torch.int32, device=device),
                       torch.zeros((batch_size, 3, num_points), device=device))
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):  # lazy initialisations (cuBLAS handles, workspaces) must not happen inside the capture
                model(*self.inputs)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.output = model(*self.inputs)

    def __call__(self, pts, mask, feats):
        k = pts.shape[0]
        for dst, src in zip(self.inputs, (pts, mask, feats)):
            dst[:k].copy_(src)  # a short last batch leaves the previous patches in the tail: computed, not used
        self.graph.replay()
        return self.output[:k]


@torch.no_grad()
def denoise_cloud(model, points, in_radius=0.05, sample_dl_patches=0.05, num_points=8192, batch_size=16, seed=0,
                  use_graph=True):
    """points (N, 3) float32 cuda -> (denoised (N, 3), mean_offset (N, 3), votes (N,))."""
    assert num_points % 128 == 0, "num_points must be a multiple of 128"
    model.eval()
    n = points.shape[0]
    centre_idx = patch_centres(points, sample_dl_patches)
    pts, mask, feats, inds = extract_patches(points, centre_idx, in_radius, num_points, seed)
    P = pts.shape[0]
    pred = torch.empty((P, 3, num_points), dtype=torch.float32, device=points.device)
    forward = _GraphedForward(model, batch_size, num_points, points.device) if use_graph and P > 2 * batch_size else model
    for i in range(0, P, batch_size):
        pred[i:i + batch_size] = forward(pts[i:i + batch_size], mask[i:i + batch_size], feats[i:i + batch_size])
    # votes: padding slots are sent to a pool of dummy points beyond N (spread out so that no segment grows long)
    flat = torch.arange(P * num_points, device=points.device).view(P, num_points)
    pool = 1 << 20
    vote_idx = torch.where(mask.bool(), inds, n + (flat & (pool - 1))).int().view(1, P * num_points // 128, 128)
    rowptr, entries = ops.build_inverse_map(vote_idx.contiguous(), n + pool)
    mean_offset, votes = ops.vote_mean(pred, rowptr, entries, n, num_points)
    return points + mean_offset, mean_offset, votes


def chamfer_ratio(clean, noisy, denoised):
    """compute_cd.py:74-90: Chamfer(clean, denoised) / Chamfer(clean, noisy)."""
    cd_noisy = ops.chamfer_l2(clean, noisy)[0]
    cd_denoised = ops.chamfer_l2(clean, denoised)[0]
    return (cd_denoised / cd_noisy).item(), cd_denoised.item(), cd_noisy.item()
