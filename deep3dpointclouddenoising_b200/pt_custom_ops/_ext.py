"""The five functions the reference's pybind module exports (u_net_arch/pt_custom_ops/_ext_src/src/bindings.cpp:8-14),
same names, argument order, dtypes, shapes and error behaviour — implemented by the sm_100a C-ABI library.

    group_points(points, idx) -> Tensor                                   bindings.cpp:8
    group_points_grad(grad_out, idx, n) -> Tensor                         bindings.cpp:9
    masked_ordered_ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample) -> [idx, idx_mask]   :11
    masked_nearest_query(query_xyz, support_xyz, query_mask, support_mask) -> [idx, idx_mask]                         :13
    masked_grid_subsampling(points, mask, nsamples, sampleDl) -> [sub_xyz, sub_mask]                                  :14
"""
from .. import ops


def group_points(points, idx):
    return ops.group_points(points, idx)


def group_points_grad(grad_out, idx, n):
    return ops.group_points_grad(grad_out, idx, n)


def masked_ordered_ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample):
    idx, idx_mask = ops.ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample)
    return [idx, idx_mask]


def masked_nearest_query(query_xyz, support_xyz, query_mask, support_mask):
    idx, idx_mask = ops.nearest_query(query_xyz, support_xyz, query_mask, support_mask)
    return [idx, idx_mask]


def masked_grid_subsampling(points, mask, nsamples, sampleDl):
    sub_xyz, sub_mask = ops.grid_subsample(points, mask, nsamples, sampleDl)
    return [sub_xyz, sub_mask]
