"""TEST INFRASTRUCTURE — float oracle for the aggregation operators: the reference's eager-PyTorch
formulas restated on CPU tensors (float32, or float64 for tight gradient checks).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Each function cites the reference lines it restates (paths relative to /root/reference/u_net_arch/).
Pinned by tests/test_oracle.py against tests/golden/aggregation_*.npz, which oracle/make_golden.py produced
by running the reference's OWN Python modules (pt_utils.py, local_aggregation_operators.py) on top of the
reference's own kernels compiled for the host (oracle/_ref/libref_emul.so).
"""
import torch


def gather(features, idx):
    """pt_utils.py:17-40 / group_points_gpu.cu:23-32: (B,C,N),(B,M,ns) -> (B,C,M,ns)."""
    B, C, N = features.shape
    _, M, ns = idx.shape
    flat = idx.reshape(B, 1, M * ns).expand(B, C, M * ns).long()
    return torch.gather(features, 2, flat).reshape(B, C, M, ns)


def clamp_idx(idx, n_support):
    """pt_utils.py:126-127."""
    idx = idx.clone()
    idx[idx > n_support] = 0
    idx[idx < 0] = 0
    return idx


def relative_positions(query_xyz, support_xyz, idx, radius=None):
    """pt_utils.py:129-133: grouped support coordinates minus the query, optionally / radius."""
    rel = gather(support_xyz.transpose(1, 2).contiguous(), idx) - query_xyz.transpose(1, 2).unsqueeze(-1)
    return rel / radius if radius is not None else rel


def feature_mask(idx_mask, query_mask):
    """local_aggregation_operators.py:171 / :490."""
    return idx_mask + (1 - query_mask[:, :, None])


def pospool(features, query_xyz, support_xyz, query_mask, idx, idx_mask, radius, reduction, embedding='xyz'):
    """local_aggregation_operators.py:140-183 (before the BN/ReLU block)."""
    B, C, _ = features.shape
    M, ns = idx.shape[1], idx.shape[2]
    idx = clamp_idx(idx, support_xyz.shape[1])
    grouped = gather(features, idx)
    rel = relative_positions(query_xyz, support_xyz, idx, radius)
    if embedding == 'xyz':
        agg = (rel.unsqueeze(1) * grouped.view(B, C // 3, 3, M, ns)).view(B, C, M, ns)
    else:
        fd = C // 6
        dim_mat = torch.pow(torch.tensor(1000.0, dtype=features.dtype), torch.arange(fd, dtype=features.dtype) / fd)
        div = (100 * rel).unsqueeze(-1) / dim_mat
        emb = torch.cat([div.sin(), div.cos()], -1).permute(0, 1, 4, 2, 3).reshape(B, C, M, ns)
        agg = grouped * emb
    if reduction == 'max':
        return agg.max(-1).values
    fm = feature_mask(idx_mask, query_mask)[:, None].to(features.dtype)
    out = (agg * fm).sum(-1)
    return out / fm.sum(-1) if reduction in ('avg', 'mean') else out


def pseudogrid(features, kernel_weights, k_points, query_xyz, support_xyz, query_mask, idx, idx_mask, extent,
               influence='linear'):
    """local_aggregation_operators.py:467-503 (before the BN/ReLU block)."""
    B, C, _ = features.shape
    M, ns = idx.shape[1], idx.shape[2]
    idx = clamp_idx(idx, support_xyz.shape[1])
    grouped = gather(features, idx)  # (B, C, M, ns)
    rel = relative_positions(query_xyz, support_xyz, idx).permute(0, 2, 3, 1).unsqueeze(3)  # (B, M, ns, 1, 3)
    sq = ((rel - k_points) ** 2).sum(-1)  # (B, M, ns, K)
    if influence == 'constant':
        w = torch.ones_like(sq)
    elif influence == 'linear':
        w = torch.clamp(1 - torch.sqrt(sq) / extent, min=0.0)
    else:
        sigma = extent * 0.3
        w = torch.exp(-sq / (2 * sigma ** 2 + 1e-9))
    w = w.permute(0, 1, 3, 2) * feature_mask(idx_mask, query_mask)[:, :, None, :].to(features.dtype)  # (B, M, K, ns)
    K = k_points.shape[0]
    weighted = torch.bmm(w.reshape(-1, K, ns), grouped.permute(0, 2, 3, 1).reshape(-1, ns, C))  # (B*M, K, C)
    return (weighted * kernel_weights).sum(1).view(B, M, C).transpose(1, 2)


def max_pool(features, idx):
    """pt_utils.py:136,202-205: gather then max over every nsample slot."""
    return gather(features, clamp_idx(idx, features.shape[2])).max(-1).values


def nearest_upsample(features, idx):
    """pt_utils.py:168,226: gather with the nearest index (B, M, 1), slot 0."""
    return gather(features, idx.clamp(min=0))[..., 0]
