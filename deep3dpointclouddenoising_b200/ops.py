"""Tensor-level host side of the C ABI: argument checks, output/workspace allocation, stream hand-off.

torch is used here for device memory and streams only; every byte of compute is in libd3d_b200.so.
Error behaviour mirrors the reference extension (u_net_arch/pt_custom_ops/_ext_src/include/utils.h:10-30):
wrong dtype / non-contiguous / CPU tensors raise RuntimeError with the reference's wording.
"""
import torch

from . import _lib

REDUCTIONS = {"sum": 0, "avg": 1, "mean": 1}
INFLUENCES = {"constant": 0, "linear": 1, "gaussian": 2}

launch_count = 0  # number of C-ABI compute calls issued (bench.py reports it as gpu_launches evidence)


def _count():
    global launch_count
    launch_count += 1


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError("CPU not supported")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {'a float' if dtype == torch.float32 else 'an int'} tensor")
    return t


def _f32(t, name):
    return _chk(t, torch.float32, name)


def _i32(t, name):
    return _chk(t, torch.int32, name)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _p(t):
    return t.data_ptr() if t is not None else None


class Arena:
    """One allocation that the OUTPUTS of the neighbourhood ops (lists, subsampled levels, inverse maps, orders) are carved
    from while it is installed (`with arena:`): a whole pyramid then is one contiguous block — handed from one step to
    the next with a single copy (neighbors.fold_pending_into_current).  Without a buffer it only measures (`need`)."""

    _current = None

    def __init__(self, nbytes=0, device=None):
        self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device) if nbytes else None
        self.off = 0
        self.need = 0

    def __enter__(self):
        self._prev, Arena._current = Arena._current, self
        return self

    def __exit__(self, *exc):
        Arena._current = self._prev

    def take(self, shape, dtype, device):
        n = int(torch.tensor([], dtype=dtype).element_size())
        for d in shape:
            n *= int(d)
        aligned = (n + 255) // 256 * 256
        self.need += aligned
        if self.buf is None or self.off + aligned > self.buf.numel() or n == 0:
            return torch.empty(shape, dtype=dtype, device=device)
        t = self.buf[self.off:self.off + n].view(dtype).view(shape)
        self.off += aligned
        return t


def _out(shape, dtype, device):
    """Output tensor of a neighbourhood op: from the installed Arena, else a fresh allocation."""
    a = Arena._current
    return torch.empty(shape, dtype=dtype, device=device) if a is None else a.take(shape, dtype, device)


# ------------------------------------------------------------------------------------------------
# neighbourhood construction
# ------------------------------------------------------------------------------------------------
def ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample, want_nvalid=False,
               want_by_support=False):
    """(idx, idx_mask[, nvalid][, idx_by_support]) — _ext.masked_ordered_ball_query (masked_ordered_ball_query.cpp:13-59).
    idx_by_support: the winners of every query in ascending support index, (distance rank << 16) | index, -1 padded."""
    L = _lib.load()
    q, s = _f32(query_xyz, "query_xyz"), _f32(support_xyz, "support_xyz")
    qm, sm = _i32(query_mask, "query_mask"), _i32(support_mask, "support_mask")
    B, M, N = q.shape[0], q.shape[1], s.shape[1]
    with torch.cuda.device(q.device):
        idx = _out((B, M, nsample), torch.int32, q.device)
        msk = _out((B, M, nsample), torch.int32, q.device)
        nv = _out((B, M), torch.int32, q.device) if want_nvalid else None
        want_by_support = want_by_support and N <= 65536
        bys = _out((B, M, nsample), torch.int32, q.device) if want_by_support else None
        ws = _ws(L.d3d_ball_query_workspace_bytes(B, M, N), q.device)
        _lib.check(L.d3d_ball_query(_p(q), _p(s), _p(qm), _p(sm), B, M, N, float(radius), int(nsample), _p(idx),
                                    _p(msk), _p(nv), _p(bys), _p(ws), ws.numel(), _stream()), "d3d_ball_query")
    _count()
    out = (idx, msk) + ((nv,) if want_nvalid else ()) + ((bys,) if want_by_support else ())
    return out


def nearest_query(query_xyz, support_xyz, query_mask, support_mask):
    """(idx, idx_mask) of shape (B, M, 1) — _ext.masked_nearest_query (masked_nearest_query.cpp:12-47)."""
    L = _lib.load()
    q, s = _f32(query_xyz, "query_xyz"), _f32(support_xyz, "support_xyz")
    qm, sm = _i32(query_mask, "query_mask"), _i32(support_mask, "support_mask")
    B, M, N = q.shape[0], q.shape[1], s.shape[1]
    with torch.cuda.device(q.device):
        idx = _out((B, M, 1), torch.int32, q.device)
        msk = _out((B, M, 1), torch.int32, q.device)
        ws = _ws(L.d3d_nearest_query_workspace_bytes(B), q.device)
        _lib.check(L.d3d_nearest_query(_p(q), _p(s), _p(qm), _p(sm), B, M, N, _p(idx), _p(msk), _p(ws), ws.numel(),
                                       _stream()), "d3d_nearest_query")
    _count()
    return idx, msk


def grid_subsample(xyz, mask, npoint, sample_dl):
    """(sub_xyz, sub_mask) — _ext.masked_grid_subsampling (masked_grid_subsampling.cpp:13-44)."""
    L = _lib.load()
    p, mk = _f32(xyz, "points"), _i32(mask, "mask")
    B, N = p.shape[0], p.shape[1]
    with torch.cuda.device(p.device):
        sub = _out((B, npoint, 3), torch.float32, p.device)
        subm = _out((B, npoint), torch.int32, p.device)
        ws = _ws(L.d3d_grid_subsample_workspace_bytes(B, N), p.device)
        _lib.check(L.d3d_grid_subsample(_p(p), _p(mk), B, N, int(npoint), float(sample_dl), _p(sub), _p(subm), _p(ws),
                                        ws.numel(), _stream()), "d3d_grid_subsample")
    _count()
    return sub, subm


# ------------------------------------------------------------------------------------------------
# gather in the reference layout
# ------------------------------------------------------------------------------------------------
def group_points(points, idx):
    """(B,C,N),(B,M,ns) -> (B,C,M,ns) — _ext.group_points (group_points.cpp:17-40)."""
    L = _lib.load()
    p, i = _f32(points, "points"), _i32(idx, "idx")
    B, C, N = p.shape
    M, ns = i.shape[1], i.shape[2]
    with torch.cuda.device(p.device):
        out = torch.empty((B, C, M, ns), dtype=torch.float32, device=p.device)
        _lib.check(L.d3d_group_points(_p(p), _p(i), B, C, N, M, ns, _p(out), _stream()), "d3d_group_points")
    _count()
    return out


def group_points_grad(grad_out, idx, n, deterministic=None):
    """(B,C,M,ns),(B,M,ns) -> (B,C,n) — _ext.group_points_grad (group_points.cpp:42-65).  deterministic=False (the default unless
    `runtime.deterministic_scatter`): the reference's own atomicAdd formulation on shared-memory planes, 7x faster;
    True: fixed summation order over an inverse map built inside the call (bit-reproducible)."""
    from .utils.config import runtime
    L = _lib.load()
    g, i = _f32(grad_out, "grad_out"), _i32(idx, "idx")
    B, C, M, ns = g.shape
    if deterministic is None:
        deterministic = bool(runtime.deterministic_scatter)
    with torch.cuda.device(g.device):
        out = torch.empty((B, C, int(n)), dtype=torch.float32, device=g.device)
        if not deterministic and int(n) * 4 <= 200 * 1024:
            _lib.check(L.d3d_group_points_grad_atomic(_p(g), _p(i), B, C, int(n), M, ns, _p(out), _stream()),
                       "d3d_group_points_grad_atomic")
        else:
            ws = _ws(L.d3d_group_points_grad_workspace_bytes(B, int(n), M, ns), g.device)
            _lib.check(L.d3d_group_points_grad(_p(g), _p(i), B, C, int(n), M, ns, _p(out), _p(ws), ws.numel(), _stream()),
                       "d3d_group_points_grad")
    _count()
    return out


def build_inverse_map(idx, n_support):
    """CSR (rowptr (B*N+1,), entries (B*M*ns,)) of idx (B, M, ns) or (B, M)."""
    L = _lib.load()
    i = _i32(idx, "idx")
    if i.dim() == 2:
        i = i.unsqueeze(-1)
    B, M, ns = i.shape
    with torch.cuda.device(i.device):
        rowptr = _out((B * n_support + 1,), torch.int32, i.device)
        entries = _out((max(B * M * ns, 1),), torch.int32, i.device)
        ws = _ws(L.d3d_inverse_map_workspace_bytes(B, n_support, M, ns), i.device)
        _lib.check(L.d3d_build_inverse_map(_p(i), B, int(n_support), M, ns, _p(rowptr), _p(entries), _p(ws),
                                           ws.numel(), _stream()), "d3d_build_inverse_map")
    _count()
    return rowptr, entries


def cm_to_cl(x):
    """(B, C, N) -> (B, N, C)."""
    L = _lib.load()
    x = _f32(x, "features")
    B, C, N = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty((B, N, C), dtype=torch.float32, device=x.device)
        _lib.check(L.d3d_cm_to_cl(_p(x), B, C, N, _p(out), _stream()), "d3d_cm_to_cl")
    _count()
    return out


def cl_to_cm(x):
    """(B, N, C) -> (B, C, N)."""
    L = _lib.load()
    x = _f32(x, "features")
    B, N, C = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty((B, C, N), dtype=torch.float32, device=x.device)
        _lib.check(L.d3d_cl_to_cm(_p(x), B, C, N, _p(out), _stream()), "d3d_cl_to_cm")
    _count()
    return out


# ------------------------------------------------------------------------------------------------
# fused aggregation (channel-last tensors)
# ------------------------------------------------------------------------------------------------
TILE_MAX_POINTS, TILE_MAX_NSAMPLE = 16384, 64  # limits of the staged-tile kernels (csrc/pospool_tiles.cu)


def spatial_order(xyz):
    """(B, N, 3) -> (B, N) int32 Morton processing order of every cloud, or None beyond the staged kernels' size limit."""
    L = _lib.load()
    p = _f32(xyz, "points")
    B, N = p.shape[0], p.shape[1]
    if N > TILE_MAX_POINTS:
        return None
    with torch.cuda.device(p.device):
        order = _out((B, N), torch.int32, p.device)
        _lib.check(L.d3d_spatial_order(_p(p), B, N, _p(order), _stream()), "d3d_spatial_order")
    _count()
    return order


def _tiles_ok(M, N, ns, C):
    return 0 < M <= TILE_MAX_POINTS and N <= TILE_MAX_POINTS and ns <= TILE_MAX_NSAMPLE and C % 4 == 0


def tile_plan(idx_by_support, nvalid, query_mask, query_order, n_support):
    """Tile plan (uint8 buffer) of one (neighbour list, query order) pair for the staged-tile PosPool kernels: union
    sizes, unions and union ranks per tile of 128 queries (d3d_pospool_tile_plan); shared by forward and backward."""
    L = _lib.load()
    bys = _i32(idx_by_support, "idx_by_support")
    B, M, ns = bys.shape
    with torch.cuda.device(bys.device):
        plan = _out((max(int(L.d3d_pospool_tile_plan_bytes(B, M, int(n_support), ns)), 16),), torch.uint8, bys.device)
        _lib.check(L.d3d_pospool_tile_plan(_p(bys), _p(_i32(nvalid, "nvalid")), _p(_i32(query_mask, "query_mask")),
                                           _p(_i32(query_order, "query_order")), B, M, int(n_support), ns, _p(plan),
                                           plan.numel(), _stream()), "d3d_pospool_tile_plan")
    _count()
    return plan


def pospool_fwd(feat_cl, query_xyz, support_xyz, idx, nvalid, query_mask, radius, reduction, query_order=None,
                idx_by_support=None, plan=None):
    """query_order (B, M) from `spatial_order` + idx_by_support from `ball_query` (+ the pair's `tile_plan`, built here
    when not given): the staged-tile tensor-core kernel; otherwise the per-query gather kernel."""
    L = _lib.load()
    f = _f32(feat_cl, "features")
    B, N, C = f.shape
    M, ns = idx.shape[1], idx.shape[2]
    with torch.cuda.device(f.device):
        out = torch.empty((B, M, C), dtype=torch.float32, device=f.device)
        xyz = (_p(_f32(query_xyz, "query_xyz")), _p(_f32(support_xyz, "support_xyz")))
        tail = (_p(_i32(nvalid, "nvalid")), _p(_i32(query_mask, "query_mask")))
        if query_order is not None and idx_by_support is not None and _tiles_ok(M, N, ns, C):
            if plan is None:
                plan = tile_plan(idx_by_support, nvalid, query_mask, query_order, N)
            _lib.check(L.d3d_pospool_tiles_fwd(_p(f), *xyz, _p(_i32(idx_by_support, "idx_by_support")), *tail,
                                               _p(_i32(query_order, "query_order")), _p(plan), B, M, N, C, ns, float(radius),
                                               REDUCTIONS[reduction], _p(out), _stream()), "d3d_pospool_tiles_fwd")
        else:
            _lib.check(L.d3d_pospool_fwd(_p(f), *xyz, _p(_i32(idx, "idx")), *tail, B, M, N, C, ns, float(radius),
                                         REDUCTIONS[reduction], _p(out), _stream()), "d3d_pospool_fwd")
    _count()
    return out


def pospool_bwd(grad_out_cl, query_xyz, support_xyz, rowptr, entries, nvalid, query_mask, n_support, nsample, radius,
                reduction, query_order=None, idx_by_support=None, plan=None, ordered=True):
    """query_order (B, M) + idx_by_support (the forward tile's inputs; + the pair's `tile_plan`, built here when not
    given): the scatter-form staged-tile kernel on the tensor cores (rowptr / entries are not read) — ordered=True adds
    the tiles' partial rows per support in a fixed order (no float atomics, bit-reproducible), ordered=False with float
    atomics like the reference; otherwise the per-support segmented reduction over the inverse map."""
    L = _lib.load()
    g = _f32(grad_out_cl, "grad_out")
    B, M, C = g.shape
    with torch.cuda.device(g.device):
        out = torch.empty((B, n_support, C), dtype=torch.float32, device=g.device)
        if query_order is not None and idx_by_support is not None and _tiles_ok(M, int(n_support), int(nsample), C):
            if plan is None:
                plan = tile_plan(idx_by_support, nvalid, query_mask, query_order, n_support)
            ws = _ws(L.d3d_pospool_scatter_bwd_workspace_bytes(B, M, int(n_support), C, int(nsample)), g.device) if ordered else None
            _lib.check(L.d3d_pospool_scatter_bwd(_p(g), _p(query_xyz), _p(support_xyz),
                                                 _p(_i32(idx_by_support, "idx_by_support")), _p(nvalid), _p(query_mask),
                                                 _p(_i32(query_order, "query_order")), _p(plan), B, M, int(n_support), C,
                                                 int(nsample), float(radius), REDUCTIONS[reduction], _p(out), _p(ws),
                                                 ws.numel() if ws is not None else 0, _stream()),
                       "d3d_pospool_scatter_bwd")
        else:
            _lib.check(L.d3d_pospool_bwd(_p(g), _p(query_xyz), _p(support_xyz), _p(rowptr), _p(entries), _p(nvalid),
                                         _p(query_mask), B, M, int(n_support), C, int(nsample), float(radius),
                                         REDUCTIONS[reduction], _p(out), _stream()), "d3d_pospool_bwd")
    _count()
    return out


def pseudogrid_fwd(feat_cl, query_xyz, support_xyz, idx, nvalid, query_mask, kpoints, weights, extent, influence,
                   precision=0):
    L = _lib.load()
    f = _f32(feat_cl, "features")
    B, N, C = f.shape
    M, ns = idx.shape[1], idx.shape[2]
    kp, w = _f32(kpoints, "K_points"), _f32(weights, "kernel_weights")
    K = kp.shape[0]
    with torch.cuda.device(f.device):
        out = torch.empty((B, M, C), dtype=torch.float32, device=f.device)
        _lib.check(L.d3d_pseudogrid_fwd(_p(f), _p(_f32(query_xyz, "query_xyz")), _p(_f32(support_xyz, "support_xyz")),
                                        _p(_i32(idx, "idx")), _p(_i32(nvalid, "nvalid")),
                                        _p(_i32(query_mask, "query_mask")), _p(kp), _p(w), B, M, N, C, ns, K,
                                        float(extent), INFLUENCES[influence], int(precision), _p(out), _stream()),
                   "d3d_pseudogrid_fwd")
    _count()
    return out


def pseudogrid_bwd(grad_out_cl, feat_cl, query_xyz, support_xyz, idx, rowptr, entries, nvalid, query_mask, kpoints,
                   weights, extent, influence, precision=0, need_feat=True, need_weights=True):
    L = _lib.load()
    g, f = _f32(grad_out_cl, "grad_out"), _f32(feat_cl, "features")
    B, M, C = g.shape
    N = f.shape[1]
    ns = idx.shape[2]
    K = kpoints.shape[0]
    with torch.cuda.device(g.device):
        gf = torch.empty((B, N, C), dtype=torch.float32, device=g.device) if need_feat else None
        gw = torch.empty((K, C), dtype=torch.float32, device=g.device) if need_weights else None
        ws = _ws(L.d3d_pseudogrid_bwd_workspace_bytes(B, M, C, K), g.device)
        _lib.check(L.d3d_pseudogrid_bwd(_p(g), _p(f), _p(query_xyz), _p(support_xyz), _p(idx), _p(rowptr), _p(entries),
                                        _p(nvalid), _p(query_mask), _p(kpoints), _p(weights), B, M, N, C, ns, K,
                                        float(extent), INFLUENCES[influence], int(precision), _p(gf), _p(gw), _p(ws), ws.numel(),
                                        _stream()), "d3d_pseudogrid_bwd")
    _count()
    return gf, gw


def gather_max_fwd(feat_cl, idx):
    L = _lib.load()
    f, i = _f32(feat_cl, "features"), _i32(idx, "idx")
    B, N, C = f.shape
    M, ns = i.shape[1], i.shape[2]
    with torch.cuda.device(f.device):
        out = torch.empty((B, M, C), dtype=torch.float32, device=f.device)
        arg = torch.empty((B, M, C), dtype=torch.uint8, device=f.device)
        _lib.check(L.d3d_gather_max_fwd(_p(f), _p(i), B, M, N, C, ns, _p(out), _p(arg), _stream()), "d3d_gather_max_fwd")
    _count()
    return out, arg


def gather_max_bwd(grad_out_cl, argslot, rowptr, entries, n_support):
    L = _lib.load()
    g = _f32(grad_out_cl, "grad_out")
    B, M, C = g.shape
    with torch.cuda.device(g.device):
        out = torch.empty((B, n_support, C), dtype=torch.float32, device=g.device)
        _lib.check(L.d3d_gather_max_bwd(_p(g), _p(argslot), _p(rowptr), _p(entries), B, M, int(n_support), C, _p(out),
                                        _stream()), "d3d_gather_max_bwd")
    _count()
    return out


def nearest_gather_fwd(feat_cl, idx):
    L = _lib.load()
    f, i = _f32(feat_cl, "features"), _i32(idx, "idx")
    B, N, C = f.shape
    M = i.shape[1]
    with torch.cuda.device(f.device):
        out = torch.empty((B, M, C), dtype=torch.float32, device=f.device)
        _lib.check(L.d3d_nearest_gather_fwd(_p(f), _p(i), B, M, N, C, _p(out), _stream()), "d3d_nearest_gather_fwd")
    _count()
    return out


def nearest_gather_bwd(grad_out_cl, rowptr, entries, n_support):
    L = _lib.load()
    g = _f32(grad_out_cl, "grad_out")
    B, M, C = g.shape
    with torch.cuda.device(g.device):
        out = torch.empty((B, n_support, C), dtype=torch.float32, device=g.device)
        _lib.check(L.d3d_nearest_gather_bwd(_p(g), _p(rowptr), _p(entries), B, M, int(n_support), C, _p(out), _stream()),
                   "d3d_nearest_gather_bwd")
    _count()
    return out


# ------------------------------------------------------------------------------------------------
# fused BatchNorm1d (+ residual) (+ ReLU), channel-major tensors
# ------------------------------------------------------------------------------------------------
_bn_workspaces = {}


def _bn_ws(device, C):
    """Zero-initialised, reused workspace of the split channel reductions: one per (device, channel count, STREAM) — two
    streams running BatchNorm layers of the same width concurrently must not share partial sums and ticket counters.
    The kernels leave the counters zero.  (During CUDA-graph capture the capturing stream is the key; replays of the
    graph are ordered among themselves.)"""
    key = (device, C, torch.cuda.current_stream(device).cuda_stream)
    ws = _bn_workspaces.get(key)
    if ws is None:
        ws = torch.zeros(_lib.load().d3d_bn_act_workspace_bytes(C), dtype=torch.uint8, device=device)
        _bn_workspaces[key] = ws
    return ws


def bn_act_fwd(x, residual, gamma, beta, running_mean, running_var, eps, momentum, training, relu):
    L = _lib.load()
    x = _f32(x, "input")
    B, C, N = x.shape
    with torch.cuda.device(x.device):
        y = torch.empty_like(x)
        mean = torch.empty((C,), dtype=torch.float32, device=x.device)
        invstd = torch.empty((C,), dtype=torch.float32, device=x.device)
        ws = _bn_ws(x.device, C)
        _lib.check(L.d3d_bn_act_fwd(_p(x), _p(residual), _p(gamma), _p(beta), _p(running_mean), _p(running_var), B, C, N,
                                    float(eps), float(momentum), int(bool(training)), int(bool(relu)), _p(y), _p(mean),
                                    _p(invstd), _p(ws), ws.numel(), _stream()), "d3d_bn_act_fwd")
    _count()
    return y, mean, invstd


def bn_act_bwd(dy, x, y, gamma, beta, mean, invstd, training, relu_mode, need_res):
    """relu_mode: 0 none, 1 ReLU mask recomputed from x (no residual), 2 mask from y."""
    L = _lib.load()
    dy = _f32(dy, "grad_out")
    B, C, N = dy.shape
    with torch.cuda.device(dy.device):
        dx = torch.empty_like(dy)
        dres = torch.empty_like(dy) if need_res else None
        dgamma = torch.empty((C,), dtype=torch.float32, device=dy.device)
        dbeta = torch.empty((C,), dtype=torch.float32, device=dy.device)
        ws = _bn_ws(dy.device, C)
        _lib.check(L.d3d_bn_act_bwd(_p(dy), _p(x), _p(y), _p(gamma), _p(beta), _p(mean), _p(invstd), B, C, N,
                                    int(bool(training)), int(relu_mode), _p(dx), _p(dres), _p(dgamma), _p(dbeta), _p(ws),
                                    ws.numel(), _stream()), "d3d_bn_act_bwd")
    _count()
    return dx, dres, dgamma, dbeta


def bn_act_cl_fwd(x_rows, residual_rows, gamma, beta, running_mean, running_var, eps, momentum, training, relu,
                  num_batches_tracked=None):
    """Channel-last variant: x_rows is a contiguous (..., C) tensor (rows = all leading dims).  num_batches_tracked:
    the module's int64 counter, incremented inside the statistics kernel in training mode."""
    if num_batches_tracked is not None and not (training and num_batches_tracked.dtype == torch.int64):
        num_batches_tracked = None
    L = _lib.load()
    x_rows = _f32(x_rows, "input")
    C = x_rows.shape[-1]
    R = x_rows.numel() // C
    with torch.cuda.device(x_rows.device):
        y = torch.empty_like(x_rows)
        mean = torch.empty((C,), dtype=torch.float32, device=x_rows.device)
        invstd = torch.empty((C,), dtype=torch.float32, device=x_rows.device)
        ws = _bn_ws(x_rows.device, C)
        _lib.check(L.d3d_bn_act_cl_fwd(_p(x_rows), _p(residual_rows), _p(gamma), _p(beta), _p(running_mean),
                                       _p(running_var), _p(num_batches_tracked), R, C, float(eps), float(momentum),
                                       int(bool(training)),
                                       int(bool(relu)), _p(y), _p(mean), _p(invstd), _p(ws), ws.numel(), _stream()),
                   "d3d_bn_act_cl_fwd")
    _count()
    return y, mean, invstd


def gemm_tf32(a0, b, a1=None, out=None, accumulate=False, want_stats=False, bias=None, residual=None, relu=False):
    """[a0 | a1] (.., K0 | K1) @ b (N, K0 + K1)^T -> (.., N) on the TF32 tensor cores (csrc/gemm.cu).
    want_stats: also the per-tile column statistics for `bn_from_stats` (the convolution feeds a BatchNorm).
    bias (N) / residual (.., N) / relu: the inference epilogue act(acc + bias + residual) (eval-mode BatchNorm folded
    into the weights by the caller)."""
    L = _lib.load()
    a0, b = _f32(a0, "rows"), _f32(b, "weight")
    K0, N = a0.shape[-1], b.shape[0]
    M = a0.numel() // K0
    K1 = 0 if a1 is None else _f32(a1, "rows").shape[-1]
    with torch.cuda.device(a0.device):
        c = out if out is not None else torch.empty(a0.shape[:-1] + (N,), dtype=torch.float32, device=a0.device)
        stats = torch.empty((L.d3d_gemm_row_tiles(M), 2, N), dtype=torch.float32, device=a0.device) if want_stats else None
        if bias is None and residual is None and not relu:
            _lib.check(L.d3d_gemm_tf32(_p(a0), _p(a1), _p(b), _p(c), M, N, K0, K1, int(bool(accumulate)), _p(stats), _stream()),
                       "d3d_gemm_tf32")
        else:
            res = _f32(residual, "residual") if residual is not None else None
            _lib.check(L.d3d_gemm_tf32_act(_p(a0), _p(a1), _p(b), _p(c), M, N, K0, K1, int(bool(accumulate)), _p(stats),
                                           _p(_f32(bias, "bias") if bias is not None else None), _p(res), int(bool(relu)),
                                           _stream()), "d3d_gemm_tf32_act")
    _count()
    return (c, stats) if want_stats else c


def wgrad_tf32(dy_rows, x_rows, into=None):
    """dW (Cout, Cin) = dy (.., Cout)^T @ x (.., Cin) over all rows, on the TF32 tensor cores (csrc/gemm.cu).
    into: a (Cout, Cin) gradient buffer to ADD the result to; returns it (or a new tensor)."""
    L = _lib.load()
    dy, x = _f32(dy_rows, "grad_out"), _f32(x_rows, "rows")
    Cout, Cin = dy.shape[-1], x.shape[-1]
    R = dy.numel() // Cout
    with torch.cuda.device(dy.device):
        out = into if into is not None else torch.empty((Cout, Cin), dtype=torch.float32, device=dy.device)
        ws = _ws(L.d3d_wgrad_workspace_bytes(R, Cout, Cin), dy.device)
        _lib.check(L.d3d_wgrad_tf32(_p(dy), _p(x), _p(out), R, Cout, Cin, int(into is not None), _p(ws), ws.numel(),
                                    _stream()), "d3d_wgrad_tf32")
    _count()
    return out


def small_linear_kind(K, N, *tensors):
    """'k' / 'n' when a (R, K) x (N, K)^T product has a tiny side the skinny kernels take (csrc/linear_small.cu), else None."""
    if not all(t is None or (t.is_cuda and t.dtype == torch.float32) for t in tensors):
        return None
    if K <= 4 and N % 4 == 0 and N <= 512:  # 512: the small-N kernel stages 64 rows of the wide side in shared memory
        return 'k'
    if N <= 4 and K % 4 == 0 and K <= 512:
        return 'n'
    return None


def linear_small(x_rows, w, bias=None, w_strides=None):
    """y (R, N) = x (R, K) . w^T (+ bias), fp32 on the CUDA cores, for K <= 4 (N % 4 == 0; w_strides = (N, stride_n, stride_k)
    reads w[n, k] at w.data_ptr + n * stride_n + k * stride_k: a transposed view without a copy) or N <= 4 (K % 4 == 0)."""
    L = _lib.load()
    x = _f32(x_rows, "x")
    R, K = x.shape
    if w_strides is None:
        N = w.shape[0]
        sn, sk = w.stride(0), w.stride(1)
    else:
        N, sn, sk = w_strides
    with torch.cuda.device(x.device):
        y = torch.empty((R, N), dtype=torch.float32, device=x.device)
        if K <= 4 and N % 4 == 0:
            _lib.check(L.d3d_linear_small_k(_p(x), _p(w), int(sn), int(sk), _p(bias), R, K, N, _p(y), _stream()),
                       "d3d_linear_small_k")
        else:
            wc = _f32(w, "w")
            _lib.check(L.d3d_linear_small_n(_p(x), _p(wc), _p(bias), R, K, N, _p(y), _stream()), "d3d_linear_small_n")
    _count()
    return y


def wgrad_small(big_rows, small_rows, out_shape, transposed, into=None):
    """sum_r big[r, n] * small[r, k] as an (Nb, Ks) matrix — or its transpose (Ks, Nb) when `transposed` — written to a
    new tensor or ADDED to `into` (the parameter's gradient buffer).  Fixed summation order."""
    L = _lib.load()
    big, small = _f32(big_rows, "big"), _f32(small_rows, "small")
    R, Nb = big.shape
    Ks = small.shape[1]
    with torch.cuda.device(big.device):
        out = into if into is not None else torch.empty(out_shape, dtype=torch.float32, device=big.device)
        sn, sk = (1, Nb) if transposed else (Ks, 1)
        ws = _ws(L.d3d_wgrad_small_workspace_bytes(Nb), big.device)
        _lib.check(L.d3d_wgrad_small(_p(big), _p(small), R, Nb, Ks, _p(out), sn, sk, 1 if into is not None else 0, _p(ws),
                                     ws.numel(), _stream()), "d3d_wgrad_small")
    _count()
    return None if into is not None else out


def gemm_ok(K0, K1, N, *tensors):
    return K0 % 4 == 0 and K1 % 4 == 0 and N % 4 == 0 and all(t is None or t.data_ptr() % 16 == 0 for t in tensors)


def bn_from_stats(x_rows, residual_rows, stats, gamma, beta, running_mean, running_var, eps, momentum, relu,
                  num_batches_tracked=None):
    """Training-mode BatchNorm (+ residual) (+ ReLU) of x_rows whose tile statistics the producing GEMM already wrote:
    finalise (Chan, fp64) + apply.  Returns (y, mean, invstd) like bn_act_cl_fwd."""
    if num_batches_tracked is not None and num_batches_tracked.dtype != torch.int64:
        num_batches_tracked = None
    L = _lib.load()
    x_rows = _f32(x_rows, "input")
    C = x_rows.shape[-1]
    R = x_rows.numel() // C
    with torch.cuda.device(x_rows.device):
        y = torch.empty_like(x_rows)
        mean = torch.empty((C,), dtype=torch.float32, device=x_rows.device)
        invstd = torch.empty((C,), dtype=torch.float32, device=x_rows.device)
        _lib.check(L.d3d_bn_finalize(_p(stats), R, C, float(eps), float(momentum), _p(running_mean), _p(running_var),
                                     _p(num_batches_tracked), _p(mean), _p(invstd), _stream()), "d3d_bn_finalize")
        _lib.check(L.d3d_bn_apply_cl(_p(x_rows), _p(residual_rows), _p(gamma), _p(beta), _p(mean), _p(invstd), R, C,
                                     int(bool(relu)), _p(y), _stream()), "d3d_bn_apply_cl")
    _count()
    return y, mean, invstd


def bn_act_cl_bwd(dy_rows, x_rows, y_rows, gamma, beta, mean, invstd, training, relu_mode, need_res,
                  grad_gamma_into=None, grad_beta_into=None):
    """grad_gamma_into / grad_beta_into: the parameters' gradient buffers; when given, the kernel ADDS into them and the
    returned dgamma / dbeta are None."""
    L = _lib.load()
    dy_rows = _f32(dy_rows, "grad_out")
    C = dy_rows.shape[-1]
    R = dy_rows.numel() // C
    with torch.cuda.device(dy_rows.device):
        dx = torch.empty_like(dy_rows)
        dres = torch.empty_like(dy_rows) if need_res else None
        in_place = grad_gamma_into is not None and grad_beta_into is not None
        dgamma = grad_gamma_into if in_place else torch.empty((C,), dtype=torch.float32, device=dy_rows.device)
        dbeta = grad_beta_into if in_place else torch.empty((C,), dtype=torch.float32, device=dy_rows.device)
        ws = _bn_ws(dy_rows.device, C)
        _lib.check(L.d3d_bn_act_cl_bwd(_p(dy_rows), _p(x_rows), _p(y_rows), _p(gamma), _p(beta), _p(mean), _p(invstd), R, C,
                                       int(bool(training)), int(relu_mode), _p(dx), _p(dres), _p(dgamma), _p(dbeta),
                                       int(in_place), _p(ws), ws.numel(), _stream()), "d3d_bn_act_cl_bwd")
    _count()
    return (dx, dres, None, None) if in_place else (dx, dres, dgamma, dbeta)


# ------------------------------------------------------------------------------------------------
# exact nearest neighbours / Chamfer distance on large clouds
# ------------------------------------------------------------------------------------------------
def nn_sqdist(query_xyz, support_xyz, want_idx=False, precise=False):
    """(M,3),(N,3) -> squared distance to the nearest support (M,), optionally its index."""
    L = _lib.load()
    q, s = _f32(query_xyz, "query_xyz"), _f32(support_xyz, "support_xyz")
    M, N = q.shape[0], s.shape[0]
    with torch.cuda.device(q.device):
        d2 = torch.empty((M,), dtype=torch.float32, device=q.device)
        idx = torch.empty((M,), dtype=torch.int32, device=q.device) if want_idx else None
        ws = _ws(L.d3d_nn_workspace_bytes(N), q.device)
        _lib.check(L.d3d_nn_sqdist(_p(q), _p(s), M, N, int(bool(precise)), _p(d2), _p(idx), _p(ws), ws.numel(), _stream()),
                   "d3d_nn_sqdist")
    _count()
    return (d2, idx) if want_idx else d2


def chamfer_l2(x, y):
    """(Nx,3),(Ny,3) -> tensor [cham_x + cham_y, cham_x, cham_y] (chamfer_distance_aux.py, L2, mean reductions)."""
    L = _lib.load()
    x, y = _f32(x, "x"), _f32(y, "y")
    with torch.cuda.device(x.device):
        out = torch.empty((3,), dtype=torch.float32, device=x.device)
        ws = _ws(L.d3d_chamfer_workspace_bytes(x.shape[0], y.shape[0]), x.device)
        _lib.check(L.d3d_chamfer_l2(_p(x), _p(y), x.shape[0], y.shape[0], _p(out), _p(ws), ws.numel(), _stream()),
                   "d3d_chamfer_l2")
    _count()
    return out


# ------------------------------------------------------------------------------------------------
# full-shape inference support
# ------------------------------------------------------------------------------------------------
def voxel_ids(points, origin, dl, nx, ny, n_cells):
    L = _lib.load()
    p = _f32(points, "points")
    with torch.cuda.device(p.device):
        ids = torch.empty((p.shape[0],), dtype=torch.int32, device=p.device)
        _lib.check(L.d3d_voxel_ids(_p(p), p.shape[0], float(origin[0]), float(origin[1]), float(origin[2]), float(dl), int(nx),
                                   int(ny), int(n_cells), _p(ids), _stream()), "d3d_voxel_ids")
    _count()
    return ids


def voxel_barycentres(points, rowptr, entries, n_cells):
    L = _lib.load()
    p = _f32(points, "points")
    with torch.cuda.device(p.device):
        bary = torch.zeros((n_cells, 3), dtype=torch.float32, device=p.device)
        counts = torch.empty((n_cells,), dtype=torch.int32, device=p.device)
        _lib.check(L.d3d_voxel_barycentres(_p(p), _p(rowptr), _p(entries), int(n_cells), _p(bary), _p(counts), _stream()),
                   "d3d_voxel_barycentres")
    _count()
    return bary, counts


def radius_patches(points, centres, radius, num_points, overflow_stride=0, smem_keys=12288):
    """(idx (P, num_points) ascending distance, -1 padded; count (P)) of the points within `radius` of every centre.
    smem_keys: candidates a block holds in shared memory (12288, or a smaller power of two >= 256: more blocks per SM when
    the balls are small); a count above it means that ball overflowed — call again with more (`radius_neighbors` does)."""
    L = _lib.load()
    p, c = _f32(points, "points"), _f32(centres, "centres")
    N, P = p.shape[0], c.shape[0]
    with torch.cuda.device(p.device):
        idx = torch.empty((P, num_points), dtype=torch.int32, device=p.device)
        cnt = torch.empty((P,), dtype=torch.int32, device=p.device)
        ws = _ws(L.d3d_radius_patches_workspace_bytes(N, P, int(overflow_stride)), p.device)
        _lib.check(L.d3d_radius_patches_tier(_p(p), N, _p(c), P, float(radius), int(num_points), int(smem_keys),
                                             int(overflow_stride), _p(idx), _p(cnt), _p(ws), ws.numel(), _stream()),
                   "d3d_radius_patches_tier")
    _count()
    return idx, cnt


def radius_neighbors(points, centres, radius, num_points):
    """radius_patches with the candidate storage sized from the data: a probe over 64 centres picks the tier (2048 keys in
    shared memory, 12288, or global scratch); a ball that still overflows re-runs the call one tier up."""
    P = centres.shape[0]
    probe = centres[:: max(P // 64, 1)][:64].contiguous()
    cmax = int(radius_patches(points, probe, radius, 1)[1].max()) if probe.shape[0] else 0
    tiers = [2048, 12288]
    k = 0 if 2 * cmax <= tiers[0] else 1
    while True:
        if k < len(tiers):
            idx, cnt = radius_patches(points, centres, radius, num_points, smem_keys=tiers[k])
            limit = tiers[k]
        else:
            stride = 1 << max(int(cmax - 1).bit_length(), 14)
            idx, cnt = radius_patches(points, centres, radius, num_points, overflow_stride=stride)
            limit = stride
        cmax = int(cnt.max()) if cnt.numel() else 0
        if cmax <= limit:
            return idx, cnt
        k += 1


def vote_mean(pred, rowptr, entries, n_points, num_points):
    L = _lib.load()
    pr = _f32(pred, "pred")
    with torch.cuda.device(pr.device):
        mean = torch.empty((n_points, 3), dtype=torch.float32, device=pr.device)
        votes = torch.empty((n_points,), dtype=torch.float32, device=pr.device)
        _lib.check(L.d3d_vote_mean(_p(pr), _p(rowptr), _p(entries), int(n_points), int(num_points), _p(mean), _p(votes),
                                   _stream()), "d3d_vote_mean")
    _count()
    return mean, votes
