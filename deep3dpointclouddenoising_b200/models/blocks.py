"""Conv1d / BatchNorm1d / ReLU stacks with the reference's module layout (nn.Sequential children '0', '1', '2', so
state-dict keys are identical: ref u_net_arch/models/backbones/resnet.py:32-45, local_aggregation_operators.py:117-123,
heads/multi_dimensional_head.py:40-59) whose forward runs every BatchNorm1d (+ following ReLU, + optional residual
add) through the fused kernel (csrc/batchnorm.cu).  With `runtime.channel_last` (default) the 1x1 convolutions run
as row-major GEMMs over (B*N, Cin) rows (cuBLAS through F.linear: same arithmetic as Conv1d with kernel_size 1), so
activations stay in the channel-last layout of the aggregation kernels and no transposition is launched."""
import contextlib

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from .. import ops
from ..fused import _is_grad_buffer, batch_norm_act, cat_channels, is_channel_last, rows_of
from ..utils.config import runtime


@contextlib.contextmanager
def _conv_math():
    """The reference's Conv1d runs under torch.backends.cudnn.allow_tf32 (default True: TF32 tensor cores); the GEMM
    form follows the same switch instead of cuBLAS's own (default False: fp32 SIMT, 10x slower on these shapes)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


_ones = {}
_wgrad_streams, _wgrad_pending = {}, set()


def _wgrad_side(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _wgrad_streams:
        _wgrad_streams[key] = torch.cuda.Stream(device=key)
    return key, _wgrad_streams[key]


def join_weight_grads():
    """Makes the current stream wait for the weight-gradient GEMMs that backward passes put on the side stream
    (runtime.wgrad_side_stream).  Call before anything reads the gradient buffers — distributed.FlatParameters.reduce()
    does; inside a CUDA-graph capture this is the join of the forked branch."""
    for key in list(_wgrad_pending):
        torch.cuda.current_stream(key).wait_stream(_wgrad_streams[key])
        _wgrad_pending.discard(key)


def wait_weight_grads(stream):
    """Makes `stream` wait for the weight gradients enqueued so far, without ending the fork (the join stays pending)."""
    for key in list(_wgrad_pending):
        stream.wait_stream(_wgrad_streams[key])


def _weight_grad_into(g2, x2, into):
    """dW accumulated into the parameter's gradient buffer.  The weight gradient of a layer is a leaf of the backward
    graph — nothing downstream waits for it before the optimiser — so it runs on a side stream, concurrently with the data
    gradients / BatchNorm / aggregation kernels of the layers below (measured on B200: these GEMMs are 0.76 ms of
    single-kernel time per step otherwise)."""
    if not (runtime.wgrad_side_stream and g2.is_cuda):
        _weight_grad(g2, x2, into)
        return
    key, side = _wgrad_side(g2.device)
    side.wait_stream(torch.cuda.current_stream())
    g2.record_stream(side)  # both die when this backward node returns: keep their memory until the side stream is done
    x2.record_stream(side)
    with torch.cuda.stream(side):
        _weight_grad(g2, x2, into)
    _wgrad_pending.add(key)


def _weight_grad(g2, x2, into=None):
    """dW = g^T x for (R, Cout), (R, Cin) with R >> Cout, Cin: the package's tensor-core weight-gradient GEMM
    (csrc/gemm.cu: both operands MN-major, rows split over the CTAs, deterministic reduction) when TF32 is the
    convolution arithmetic; otherwise torch / cuBLAS:  One cuBLAS GEMM with a 131072-long reduction and a
    72 x 72 result runs on a handful of CTAs (114 us measured on B200, tools/wgrad_probe.py); split into R / 1024
    independent row chunks (bmm) plus a reduction it takes 32 us.
    into: a (Cout, Cin) gradient buffer to ADD the result to (runtime.grads_in_place); returns None then."""
    R = g2.shape[0]
    if (runtime.own_wgrad and g2.is_cuda and torch.backends.cudnn.allow_tf32 and R >= runtime.own_wgrad_min_rows
            and ops.gemm_ok(g2.shape[1], 0, x2.shape[1], g2, x2, into) and g2.is_contiguous() and x2.is_contiguous()):
        out = ops.wgrad_tf32(g2, x2, into=into)
        return None if into is not None else out
    S = min(R // 1024, 128)
    if R < 16384 or R % S:
        if into is None:
            return g2.t() @ x2
        into.addmm_(g2.t(), x2)
        return None
    partial = torch.bmm(g2.view(S, R // S, -1).transpose(1, 2), x2.view(S, R // S, -1))
    if into is None:
        return partial.sum(0)
    key = (S, g2.device)
    if key not in _ones:
        _ones[key] = torch.ones((1, S), dtype=g2.dtype, device=g2.device)
    # the sum over the chunks, accumulated into the buffer in ONE launch — in full fp32: this call runs inside
    # _conv_math(), where matmul TF32 follows cudnn.allow_tf32, and TF32 would round every fp32 partial to 10 mantissa
    # bits before the sum (the out-of-place path sums with partial.sum(0) in fp32)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        into.view(1, -1).addmm_(_ones[key], partial.view(S, -1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    return None


def _own_gemm(k0, k1, n, *tensors):
    """The TF32 tensor-core GEMM of this package (csrc/gemm.cu) serves a convolution when TF32 is the convolution
    arithmetic (torch.backends.cudnn.allow_tf32, the reference's default) and TMA's alignment rules hold; otherwise
    (exact-fp32 parity runs, 3-channel inputs / outputs) the GEMM goes to cuBLAS."""
    return (runtime.own_gemm and torch.backends.cudnn.allow_tf32 and tensors[0].is_cuda
            and ops.gemm_ok(k0, k1, n, *tensors))


def _transposed(w):
    """(Cout, Cin) -> contiguous (Cin, Cout) with the package's transposition kernel (B operand of the data gradient)."""
    return ops.cm_to_cl(w.contiguous().view(1, w.shape[0], w.shape[1])).view(w.shape[1], w.shape[0])


class PointwiseConvRows(Function):
    """Conv1d(kernel_size=1) on channel-last rows: (B, N, Cin) x (Cout, Cin, 1) -> (B, N, Cout).  Forward and data
    gradient on the package's TF32 tensor-core GEMM (weight gradient: batched split-K GEMM).  want_stats: a second,
    non-differentiable output carries the per-tile column statistics for the BatchNorm that follows (ops.bn_from_stats)."""

    @staticmethod
    def forward(ctx, rows, weight, bias, want_stats=False):
        w = weight.squeeze(-1)
        ctx.save_for_backward(rows, w)
        ctx.has_bias = bias is not None
        ctx.wparam = weight  # the Parameter: its .grad buffer is written in place under runtime.grads_in_place
        ctx.set_materialize_grads(False)  # no zero tensor for the (non-differentiable) statistics output
        ctx.small = (ops.small_linear_kind(w.shape[1], w.shape[0], rows, w, bias)
                     if runtime.own_gemm and runtime.own_small_linear and rows.is_contiguous() and w.is_contiguous() else None)
        if ctx.small:  # 3-channel input / output layers: streaming kernels on the CUDA cores (csrc/linear_small.cu)
            y = ops.linear_small(rows.reshape(-1, rows.shape[-1]), w, bias).view(*rows.shape[:-1], w.shape[0])
            return (y, None) if want_stats else y
        if bias is None and _own_gemm(w.shape[1], 0, w.shape[0], rows, w) and rows.is_contiguous():
            out = ops.gemm_tf32(rows, w, want_stats=want_stats)
            if want_stats:
                ctx.mark_non_differentiable(out[1])
            return out
        with _conv_math():
            y = F.linear(rows, w, bias)
        return (y, None) if want_stats else y

    @staticmethod
    def backward(ctx, grad, _grad_stats=None):
        rows, w = ctx.saved_tensors
        if grad is None:
            return None, None, None, None
        g2 = grad.contiguous().view(-1, grad.shape[-1])
        d_rows = d_w = d_b = None
        if ctx.small:
            x2 = rows.reshape(-1, rows.shape[-1])
            N, K = w.shape
            if ctx.needs_input_grad[0]:
                if ctx.small == 'n':  # dx = dy (R, N <= 4) . W (N, K): the small-K kernel over a strided view of W
                    d_rows = ops.linear_small(g2, w, None, w_strides=(K, 1, K)).view_as(rows)
                else:                 # dx (R, K <= 4) = dy (R, N) . W: the small-N kernel with W^T (K, N)
                    d_rows = ops.linear_small(g2, w.t().contiguous(), None).view_as(rows)
            if ctx.needs_input_grad[1]:
                into = ctx.wparam.grad.view(w.shape) if runtime.grads_in_place and _is_grad_buffer(ctx.wparam) else None
                if ctx.small == 'n':
                    d_w = ops.wgrad_small(x2, g2, (N, K), True, into=into)
                else:
                    d_w = ops.wgrad_small(g2, x2, (N, K), False, into=into)
                d_w = d_w.unsqueeze(-1) if d_w is not None else None
            if ctx.has_bias and ctx.needs_input_grad[2]:
                d_b = g2.sum(0)
            return d_rows, d_w, d_b, None
        if ctx.needs_input_grad[0] and _own_gemm(w.shape[0], 0, w.shape[1], g2, w):
            d_rows = ops.gemm_tf32(g2, _transposed(w)).view_as(rows)
        with _conv_math():
            if ctx.needs_input_grad[0] and d_rows is None:
                d_rows = (g2 @ w).view_as(rows)
            if ctx.needs_input_grad[1]:
                into = ctx.wparam.grad.view(w.shape) if runtime.grads_in_place and _is_grad_buffer(ctx.wparam) else None
                if into is not None:
                    _weight_grad_into(g2, rows.reshape(-1, rows.shape[-1]), into)
                else:
                    d_w = _weight_grad(g2, rows.reshape(-1, rows.shape[-1])).unsqueeze(-1)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_b = g2.sum(0)
        return d_rows, d_w, d_b, None


class PointwiseConvCatRows(Function):
    """Conv1d(kernel_size=1) applied to the channel concatenation of several row tensors WITHOUT materialising it:
    y = sum_i rows_i @ W[:, slice_i]^T.  The decoder's skip concatenations (multi_dimensional_head.py:36) then cost no
    copy forward and hand each branch a contiguous gradient backward (a torch.cat would give strided slices that the
    next kernel has to re-pack)."""

    @staticmethod
    def forward(ctx, weight, bias, want_stats, *rows_list):
        w = weight.squeeze(-1)
        ctx.save_for_backward(w, *rows_list)
        ctx.has_bias = bias is not None
        ctx.set_materialize_grads(False)
        if (bias is None and len(rows_list) == 2 and all(r.is_contiguous() for r in rows_list)
                and _own_gemm(rows_list[0].shape[-1], rows_list[1].shape[-1], w.shape[0], rows_list[0], rows_list[1], w)):
            out = ops.gemm_tf32(rows_list[0], w, a1=rows_list[1], want_stats=want_stats)  # both K segments in one kernel
            if want_stats:
                ctx.mark_non_differentiable(out[1])
            return out
        with _conv_math():
            c0 = rows_list[0].shape[-1]
            y = F.linear(rows_list[0], w[:, :c0], bias)
            y2 = y.view(-1, y.shape[-1])
            for rows in rows_list[1:]:
                c1 = c0 + rows.shape[-1]
                y2.addmm_(rows.reshape(-1, rows.shape[-1]), w[:, c0:c1].t())
                c0 = c1
        return (y, None) if want_stats else y

    @staticmethod
    def backward(ctx, grad, _grad_stats=None):
        w, *rows_list = ctx.saved_tensors
        if grad is None:
            return (None,) * (3 + len(rows_list))
        g2 = grad.contiguous().view(-1, grad.shape[-1])
        d_rows, d_w, c0 = [], [], 0
        own = _own_gemm(w.shape[0], 0, 4, g2, w) and all(r.shape[-1] % 4 == 0 for r in rows_list)
        w_t = _transposed(w) if own and any(ctx.needs_input_grad[3:]) else None  # (Cin total, Cout): row blocks per input
        with _conv_math():
            for i, rows in enumerate(rows_list):
                c1 = c0 + rows.shape[-1]
                if not ctx.needs_input_grad[3 + i]:
                    d_rows.append(None)
                elif w_t is not None:
                    d_rows.append(ops.gemm_tf32(g2, w_t[c0:c1]).view_as(rows))
                else:
                    d_rows.append((g2 @ w[:, c0:c1]).view_as(rows))
                if ctx.needs_input_grad[0]:
                    d_w.append(_weight_grad(g2, rows.reshape(-1, rows.shape[-1])))
                c0 = c1
        d_weight = torch.cat(d_w, 1).unsqueeze(-1) if ctx.needs_input_grad[0] else None
        d_b = g2.sum(0) if ctx.has_bias and ctx.needs_input_grad[1] else None
        return (d_weight, d_b, None, *d_rows)


def _is_pointwise(m):
    return (isinstance(m, nn.Conv1d) and m.kernel_size == (1,) and m.stride == (1,) and m.padding == (0,)
            and m.dilation == (1,) and m.groups == 1 and m.padding_mode == 'zeros')


class FusedSequential(nn.Sequential):
    def forward(self, x, residual=None, final_relu=False):
        """x: a (B, C, N) tensor, or a list of them standing for their channel concatenation."""
        mods = list(self._modules.values())
        parts = list(x) if isinstance(x, (list, tuple)) else None
        on_gpu = parts[0].is_cuda if parts is not None else x.is_cuda
        if not on_gpu and not runtime.cpu_modules:
            # the product has no CPU path; the CPU oracle (oracle/cpu_model.py: tests, bench's CPU arm) opts in explicitly
            raise RuntimeError("CPU not supported: the convolution / BatchNorm blocks run on CUDA tensors only "
                               "(utils.config.runtime.cpu_modules = True lets the CPU oracle drive the module tree)")
        use_fused = on_gpu and runtime.fused_batchnorm

        def feeds_training_bn(k):
            """The convolution at position k is followed by a BatchNorm in training mode: its GEMM emits the tile
            statistics and the BatchNorm skips its own statistics pass."""
            nxt = mods[k + 1] if k + 1 < len(mods) else None
            return (use_fused and isinstance(nxt, nn.BatchNorm1d) and (nxt.training or nxt.running_mean is None)
                    and nxt.momentum is not None)

        def folded_eval(k):
            """Inference: the convolution at position k is followed by a BatchNorm in eval mode — returns (weights scaled
            per output channel, shift, relu after it?, is it the block's last?) so that conv + BN (+ residual) (+ ReLU) run
            as ONE GEMM with the epilogue of d3d_gemm_tf32_act; None otherwise.  The folded weights are cached per block
            and recomputed when a parameter or running statistic changes."""
            nxt = mods[k + 1] if k + 1 < len(mods) else None
            conv = mods[k]
            if not (use_fused and runtime.channel_last and runtime.own_gemm and runtime.fold_eval_batchnorm
                    and not torch.is_grad_enabled() and isinstance(nxt, nn.BatchNorm1d) and not nxt.training
                    and nxt.running_mean is not None and torch.backends.cudnn.allow_tf32):
                return None
            w = conv.weight.squeeze(-1)
            if not ops.gemm_ok(w.shape[1], 0, w.shape[0], w):
                return None
            srcs = [t for t in (conv.weight, conv.bias, nxt.weight, nxt.bias, nxt.running_mean, nxt.running_var) if t is not None]
            key = tuple((t.data_ptr(), t._version) for t in srcs)
            cache = self.__dict__.setdefault("_folded", {})
            hit = cache.get(k)
            if hit is None or hit[0] != key:
                scale = torch.rsqrt(nxt.running_var + nxt.eps)
                if nxt.weight is not None:
                    scale = scale * nxt.weight
                shift = -nxt.running_mean * scale
                if nxt.bias is not None:
                    shift = shift + nxt.bias
                if conv.bias is not None:
                    shift = shift + conv.bias * scale
                hit = cache[k] = (key, (w * scale[:, None]).contiguous(), shift.contiguous())
            relu_next = k + 2 < len(mods) and isinstance(mods[k + 2], nn.ReLU)
            return hit[1], hit[2], relu_next, k + (3 if relu_next else 2) == len(mods)

        stats = None  # tile statistics of x from the GEMM that produced it (consumed by the BatchNorm right after)
        if parts is not None:
            fold = folded_eval(0) if (on_gpu and len(parts) == 2 and _is_pointwise(mods[0])
                                      and all(is_channel_last(t) for t in parts)) else None
            if fold is not None and fold[3] and residual is not None and not is_channel_last(residual):
                fold = None
            if fold is not None and all(t.shape[1] % 4 == 0 for t in parts):
                w2, shift, relu_next, is_last = fold
                res = rows_of(residual) if (is_last and residual is not None) else None
                x = ops.gemm_tf32(rows_of(parts[0]), w2, a1=rows_of(parts[1]), bias=shift, residual=res,
                                  relu=relu_next or (is_last and final_relu)).permute(0, 2, 1)
                if is_last:
                    residual, final_relu = None, False
                mods = mods[(3 if relu_next else 2):]
            elif on_gpu and runtime.channel_last and _is_pointwise(mods[0]) and all(is_channel_last(t) for t in parts):
                want = feeds_training_bn(0)
                x = PointwiseConvCatRows.apply(mods[0].weight, mods[0].bias, want, *[t.permute(0, 2, 1) for t in parts])
                if want:
                    x, stats = x
                x = x.permute(0, 2, 1)
                mods = mods[1:]
            else:
                x = cat_channels(parts)
        use_rows = x.is_cuda and runtime.channel_last and x.dim() == 3
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.BatchNorm1d) and use_fused and x.dim() == 3:
                next_is_relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                is_last = i + (2 if next_is_relu else 1) == len(mods)
                res = residual if is_last else None
                x = batch_norm_act(m, x, relu=next_is_relu or (is_last and final_relu), residual=res, stats=stats)
                stats = None
                if is_last:
                    residual, final_relu = None, False
                i += 2 if next_is_relu else 1
                continue
            stats = None
            fold = folded_eval(i) if (use_rows and _is_pointwise(m) and is_channel_last(x)) else None
            if fold is not None:
                w2, shift, relu_next, is_last = fold
                want_res = is_last and residual is not None
                if not want_res or (is_channel_last(residual) and residual.shape == (x.shape[0], w2.shape[0], x.shape[2])):
                    res = rows_of(residual) if want_res else None
                    x = ops.gemm_tf32(rows_of(x), w2, bias=shift, residual=res,
                                      relu=relu_next or (is_last and final_relu)).permute(0, 2, 1)
                    if is_last:
                        residual, final_relu = None, False
                    i += 3 if relu_next else 2
                    continue
            if use_rows and _is_pointwise(m):
                want = feeds_training_bn(i)
                x = PointwiseConvRows.apply(rows_of(x), m.weight, m.bias, want)
                if want:
                    x, stats = x
                x = x.permute(0, 2, 1)
            else:
                x = m(x)
            i += 1
        if residual is not None:
            x = x + residual
        if final_relu:
            x = x.relu()
        return x


def conv_bn(cin, cout, momentum=0.1, relu=True, conv=True):
    layers = [nn.Conv1d(cin, cout, kernel_size=1, bias=False)] if conv else []
    layers.append(nn.BatchNorm1d(cout, momentum=momentum))
    if relu:
        layers.append(nn.ReLU(inplace=True))
    return FusedSequential(*layers)
