"""Autograd Functions over the fused kernels.  Inputs and outputs have the reference's LOGICAL shape
(B, C, N), so the callers — the modules mirroring pt_utils.py and local_aggregation_operators.py — keep the
reference API.  The kernels work on channel-last rows (B, N, C):

* a channel-major contiguous input is transposed inside (d3d_cm_to_cl / d3d_cl_to_cm);
* a "channel-last view" — the (B, C, N) permutation of a contiguous (B, N, C) buffer — is used in place, and
  with `runtime.channel_last` (default) outputs are returned as such views.  The U-Net then never leaves the
  row layout: aggregations, fused BatchNorm (d3d_bn_act_cl_*) and the 1x1 convolutions (row-major GEMMs,
  models/blocks.py) hand rows to each other and the step has no transposition kernel at all.

Differentiable w.r.t. features (and PseudoGrid's kernel_weights) only, like the reference
(pt_utils.py:43-62: GroupingOperation returns a gradient for features, None for idx; coordinates never
receive gradients because the grouped xyz come from non-differentiable index ops on leaf tensors).
"""
import torch
from torch.autograd import Function

from . import neighbors as _neighbors
from . import ops
from .utils.config import runtime


# ---- backward stage markers (data-parallel gradient overlap, distributed.FlatParameters.overlap_with_backward) ----------
_stage_callback = [None]


def set_stage_callback(fn):
    """fn(tag) is called during backward when the gradient has passed the marker `tag` (None switches it off)."""
    _stage_callback[0] = fn


class _StageMarker(Function):
    @staticmethod
    def forward(ctx, x, tag):
        ctx.tag = tag
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad):
        if _stage_callback[0] is not None:
            _stage_callback[0](ctx.tag)
        return grad, None


def stage_marker(x, tag):
    """Identity in the forward pass; in the backward pass it reports that everything computed FROM x in the forward pass
    has finished its backward.  Free when no callback is registered (single-GPU runs skip the autograd node entirely)."""
    if _stage_callback[0] is None or not torch.is_grad_enabled() or not x.requires_grad:
        return x
    return _StageMarker.apply(x, tag)


def stage_modules(model):
    """tag -> the first module whose parameters are final once the marker `tag` fires (markers: ResNet._forward)."""
    bb = getattr(model, "backbone", model)
    out = {}
    for tag in ("layer2", "layer3", "layer4"):
        if hasattr(bb, tag):
            out[tag] = getattr(bb, tag)
    if hasattr(model, "segmentation_head"):
        out["head"] = model.segmentation_head
    return out


def _is_grad_buffer(p):
    """A parameter whose gradient buffer exists and can be written by a kernel (runtime.grads_in_place)."""
    g = getattr(p, "grad", None)
    return g is not None and g.is_contiguous() and g.dtype == torch.float32 and g.shape == p.shape


def is_channel_last(t):
    """True for the (B, C, N) view of a contiguous (B, N, C) buffer."""
    return t.dim() == 3 and not t.is_contiguous() and t.permute(0, 2, 1).is_contiguous()


def _rows(t):
    """Contiguous (B, N, C) rows of a logical (B, C, N) tensor: no copy for channel-last views, a row-wise copy for
    channel slices of one (the gradient of a skip concatenation), a transposition for channel-major tensors."""
    if t.dim() == 3 and t.stride(1) == 1 and t.shape[1] > 1:
        return t.permute(0, 2, 1).contiguous()
    return ops.cm_to_cl(t.contiguous())


def _logical(rows, channel_last):
    """Logical (B, C, N) tensor over (B, N, C) rows: a view, or a channel-major copy."""
    return rows.permute(0, 2, 1) if channel_last else ops.cl_to_cm(rows)


def rows_of(t):
    """Differentiable (B, N, C) rows of a logical (B, C, N) tensor, for torch ops (1x1 convolutions as GEMMs)."""
    return t.permute(0, 2, 1) if is_channel_last(t) else t.transpose(1, 2).contiguous()


def cat_channels(tensors):
    """torch.cat(tensors, 1) that keeps channel-last views channel-last (a plain cat would re-layout)."""
    if all(is_channel_last(t) for t in tensors):
        return torch.cat([t.permute(0, 2, 1) for t in tensors], 2).permute(0, 2, 1)
    return torch.cat(tensors, 1)


class PosPoolFunction(Function):
    """local_aggregation_operators.py:140-147,165-183 with position_embedding='xyz', reduction sum/avg."""

    @staticmethod
    def forward(ctx, features, query_xyz, support_xyz, query_mask, nbr, radius, reduction):
        feat_cl, ctx.in_cl = _rows(features), is_channel_last(features)
        # staged tiles: self queries (where they are the faster kernel) with 'avg' (the bilinear split keeps the fp32
        # parity tolerance of a mean; a plain sum of nsample terms is nsample times larger in absolute terms)
        # (strided lists: measured faster from the 512 <- 2048 level on, a tie before)
        staged = runtime.staged_tiles == 'always' or (
            runtime.staged_tiles and reduction != 'sum'
            and (query_xyz.shape[1] == support_xyz.shape[1] or support_xyz.shape[1] <= 2048))
        plan = nbr.tile_plan(query_xyz, query_mask) if staged else None
        out_cl = ops.pospool_fwd(feat_cl, query_xyz, support_xyz, nbr.idx, nbr.nvalid, query_mask, radius, reduction,
                                 query_order=_neighbors.spatial_order(query_xyz) if plan is not None else None,
                                 idx_by_support=nbr.by_support, plan=plan)
        ctx.nbr, ctx.radius, ctx.reduction = nbr, radius, reduction
        ctx.save_for_backward(query_xyz, support_xyz, query_mask)
        return _logical(out_cl, runtime.channel_last)

    @staticmethod
    def backward(ctx, grad_out):
        query_xyz, support_xyz, query_mask = ctx.saved_tensors
        nbr = ctx.nbr
        g_cl = _rows(grad_out)
        # 'scatter': the forward tile transposed on the tensor cores, float atomics across tiles (faster at every level, also
        # for strided lists where the forward tiles are not); 'ordered': the same tiles + a fixed-order second pass (no float
        # atomics, bit-reproducible); False: segmented reduction over the inverse map
        mode = runtime.staged_tiles_backward
        plan = (nbr.tile_plan(query_xyz, query_mask)
                if mode in ('scatter', 'ordered') and ctx.reduction != 'sum' and nbr.by_support is not None else None)
        if plan is not None:
            gf_cl = ops.pospool_bwd(g_cl, query_xyz, support_xyz, None, None, nbr.nvalid, query_mask, nbr.n_support,
                                    nbr.nsample, ctx.radius, ctx.reduction, query_order=_neighbors.spatial_order(query_xyz),
                                    idx_by_support=nbr.by_support, plan=plan, ordered=mode == 'ordered')
        else:
            rowptr, entries = nbr.csr()
            gf_cl = ops.pospool_bwd(g_cl, query_xyz, support_xyz, rowptr, entries, nbr.nvalid, query_mask, nbr.n_support,
                                    nbr.nsample, ctx.radius, ctx.reduction)
        return _logical(gf_cl, ctx.in_cl), None, None, None, None, None, None


class PseudoGridFunction(Function):
    """local_aggregation_operators.py:467-503 (depthwise KPConv aggregation, convolution_mode='sum')."""

    @staticmethod
    def forward(ctx, features, kernel_weights, query_xyz, support_xyz, query_mask, nbr, k_points, extent, influence,
                precision):
        feat_cl, ctx.in_cl = _rows(features), is_channel_last(features)
        w = kernel_weights.contiguous()
        out_cl = ops.pseudogrid_fwd(feat_cl, query_xyz, support_xyz, nbr.idx, nbr.nvalid, query_mask, k_points, w,
                                    extent, influence, precision)
        ctx.nbr, ctx.extent, ctx.influence, ctx.precision = nbr, extent, influence, precision
        ctx.save_for_backward(feat_cl, w, query_xyz, support_xyz, query_mask, k_points)
        return _logical(out_cl, runtime.channel_last)

    @staticmethod
    def backward(ctx, grad_out):
        feat_cl, w, query_xyz, support_xyz, query_mask, k_points = ctx.saved_tensors
        nbr = ctx.nbr
        rowptr, entries = nbr.csr()
        g_cl = _rows(grad_out)
        gf_cl, gw = ops.pseudogrid_bwd(g_cl, feat_cl, query_xyz, support_xyz, nbr.idx, rowptr, entries, nbr.nvalid,
                                       query_mask, k_points, w, ctx.extent, ctx.influence, ctx.precision,
                                       need_feat=ctx.needs_input_grad[0], need_weights=ctx.needs_input_grad[1])
        gf = _logical(gf_cl, ctx.in_cl) if gf_cl is not None else None
        return gf, gw, None, None, None, None, None, None, None, None


class GatherMaxFunction(Function):
    """grouping_operation + F.max_pool2d over the nsample axis (pt_utils.py:136,202-205)."""

    @staticmethod
    def forward(ctx, features, nbr):
        feat_cl, ctx.in_cl = _rows(features), is_channel_last(features)
        out_cl, arg = ops.gather_max_fwd(feat_cl, nbr.idx)
        ctx.nbr = nbr
        ctx.save_for_backward(arg)
        return _logical(out_cl, runtime.channel_last)

    @staticmethod
    def backward(ctx, grad_out):
        (arg,) = ctx.saved_tensors
        rowptr, entries = ctx.nbr.csr()
        g_cl = _rows(grad_out)
        gf_cl = ops.gather_max_bwd(g_cl, arg, rowptr, entries, ctx.nbr.n_support)
        return _logical(gf_cl, ctx.in_cl), None


class NearestGatherFunction(Function):
    """grouping_operation with the nearest-neighbour index, then [..., 0] (pt_utils.py:168,226)."""

    @staticmethod
    def forward(ctx, features, nbr):
        feat_cl, ctx.in_cl = _rows(features), is_channel_last(features)
        out_cl = ops.nearest_gather_fwd(feat_cl, nbr.idx.view(nbr.idx.shape[0], nbr.idx.shape[1]))
        ctx.nbr = nbr
        return _logical(out_cl, runtime.channel_last)

    @staticmethod
    def backward(ctx, grad_out):
        rowptr, entries = ctx.nbr.csr()
        g_cl = _rows(grad_out)
        gf_cl = ops.nearest_gather_bwd(g_cl, rowptr, entries, ctx.nbr.n_support)
        return _logical(gf_cl, ctx.in_cl), None


class BatchNormActFunction(Function):
    """y = act(BatchNorm1d(x) [+ residual]) in one statistics pass and one apply pass (csrc/batchnorm.cu)."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, eps, momentum, training, relu, counter=None,
                stats=None):
        ctx.training, ctx.has_res = training, residual is not None
        ctx.wparam, ctx.bparam = weight, bias  # the Parameter objects (their .grad buffers, see runtime.grads_in_place)
        # the channel-last kernels need 16-byte aligned rows and parameters (d3d_bn_act_cl_* reject anything else): a view
        # at an odd storage offset takes the channel-major kernels instead
        aligned = all(t is None or t.data_ptr() % 16 == 0 for t in (x, residual, weight, bias))
        ctx.rows = is_channel_last(x) and x.shape[1] % 4 == 0 and aligned
        ctx.relu_mode = 0 if not relu else (2 if residual is not None else 1)
        if ctx.rows:  # channel-last view in, channel-last view out: no layout change anywhere
            xr = x.permute(0, 2, 1)
            rr = _rows(residual) if residual is not None else None
            if stats is not None and training:  # statistics came out of the producing GEMM's epilogue: finalise + apply
                y, mean, invstd = ops.bn_from_stats(xr, rr, stats, weight, bias, running_mean, running_var, eps, momentum,
                                                    relu, counter)
            else:
                y, mean, invstd = ops.bn_act_cl_fwd(xr, rr, weight, bias, running_mean, running_var, eps, momentum,
                                                    training, relu, counter)  # num_batches_tracked rides on the kernel
            ctx.save_for_backward(xr, y if ctx.relu_mode == 2 else None, weight, bias, mean, invstd)
            return y.permute(0, 2, 1)
        if counter is not None:
            counter.add_(1)
        x = x.contiguous()
        res = residual.contiguous() if residual is not None else None
        y, mean, invstd = ops.bn_act_fwd(x, res, weight, bias, running_mean, running_var, eps, momentum, training, relu)
        # the ReLU mask is recomputed from x in backward unless a residual was added (then y itself is needed)
        ctx.relu_mode = 0 if not relu else (2 if residual is not None else 1)
        ctx.save_for_backward(x, y if ctx.relu_mode == 2 else None, weight, bias, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, grad_out):
        x, y, weight, bias, mean, invstd = ctx.saved_tensors
        need_res = ctx.has_res and ctx.needs_input_grad[1]
        if ctx.rows:
            into = (None, None)
            if runtime.grads_in_place and weight is not None and bias is not None and _is_grad_buffer(ctx.wparam) \
                    and _is_grad_buffer(ctx.bparam):
                into = (ctx.wparam.grad, ctx.bparam.grad)  # the kernel adds into the gradient buffers: no add kernels
            dx, dres, dgamma, dbeta = ops.bn_act_cl_bwd(_rows(grad_out), x, y, weight, bias, mean, invstd, ctx.training,
                                                        ctx.relu_mode, need_res, *into)
            dx = dx.permute(0, 2, 1)
            dres = dres.permute(0, 2, 1) if dres is not None else None
        else:
            dx, dres, dgamma, dbeta = ops.bn_act_bwd(grad_out.contiguous(), x, y, weight, bias, mean, invstd,
                                                     ctx.training, ctx.relu_mode, need_res)
        return (dx, dres, dgamma if weight is not None else None, dbeta if bias is not None else None, None, None, None,
                None, None, None, None, None)  # dgamma / dbeta are None when they were added in place


def batch_norm_act(bn, x, relu, residual=None, stats=None):
    """Applies nn.BatchNorm1d module `bn` (its parameters, buffers, momentum, eps, training flag) with the fused kernel."""
    training = bn.training or bn.running_mean is None
    counter = bn.num_batches_tracked if (bn.training and bn.num_batches_tracked is not None) else None
    if bn.momentum is None:  # cumulative moving average: the factor needs the updated count on the host
        if counter is not None:
            counter.add_(1)
            counter = None
        momentum = 1.0 / float(max(int(bn.num_batches_tracked), 1))
    else:
        momentum = bn.momentum
    return BatchNormActFunction.apply(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, momentum,
                                      training, relu, counter, stats)
