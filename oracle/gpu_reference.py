"""TEST / BASELINE INFRASTRUCTURE — the reference's GPU path on the same B200 ("the kernel to beat").

One training step exactly as the reference runs it on a GPU (u_net_arch/pt_custom_ops/pt_utils.py:122-148 and
models/local_aggregation_operators.py:140-183 / 467-503: ball query per call, materialised (B, C, npoint, nsample)
gathers, eager PyTorch arithmetic, atomicAdd scatter in backward; cuDNN Conv1d / BatchNorm1d / ReLU modules):

* the five `_ext` functions are the reference's OWN CUDA kernels, rebuilt unmodified for sm_100a
  (oracle/_ref/libref_cuda.so through oracle/cuda_ref.py);
* the Python glue is the oracle port of the reference's call sequence (oracle/cpu_model.py, oracle/aggregation_ref.py,
  pinned to the reference's own model by tests/test_model_golden.py) — the reference's .py files cannot travel to the
  GPU box;
* the model object runs with this package's fused kernels switched OFF (stock torch modules).

Used by bench.py's `reference_gpu` leg and tools/time_reference_cuda.py only.
"""
import torch

from . import aggregation_ref as agg
from . import cuda_ref
from .cpu_model import CpuUNet


class _RefGrouping(torch.autograd.Function):
    """pt_utils.py:17-62 over the reference kernels (group_points_gpu.cu:13-80, atomicAdd backward)."""

    @staticmethod
    def forward(ctx, features, idx, R):
        ctx.idx, ctx.n, ctx.R = idx, features.shape[2], R
        return R.group_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, grad_out):
        return ctx.R.group_points_grad(grad_out.contiguous(), ctx.idx, ctx.n), None, None


class RefGpuUNet(CpuUNet):
    """CpuUNet's call sequence on CUDA tensors with the reference's CUDA kernels for every index / gather op."""

    def __init__(self, model):
        super().__init__(model, index_ops=object())
        self.R = cuda_ref.RefCuda()

    def ball_query(self, q, s, qm, sm, radius, ns):
        return self.R.ball_query(q.contiguous(), s.contiguous(), qm.contiguous(), sm.contiguous(), float(radius), int(ns))

    def max_pool(self, mp, xyz, mask, feats):
        sub, subm = self.R.grid_subsampling(xyz.contiguous(), mask.contiguous(), int(mp.npoint), float(mp.sampleDl))
        idx, _ = self.ball_query(sub, xyz, subm, mask, mp.radius, mp.nsample)
        return sub, subm, agg.max_pool(feats, idx)

    def head(self, end):
        hd = self.model.segmentation_head
        feats = end["res5"][2]
        for level in range(4):
            fine, coarse = end[f"res{4 - level}"], end[f"res{5 - level}"]
            nidx, _ = self.R.nearest_query(fine[0].contiguous(), coarse[0].contiguous(), fine[1].contiguous(), coarse[1].contiguous())
            feats = agg.nearest_upsample(feats, nidx)
            feats = torch.cat([feats, fine[2]], 1)
            feats = getattr(hd, f"up_conv{level}")(feats)
        return hd.head(feats)

    def forward(self, xyz, mask, feats):
        old = agg.gather
        agg.gather = lambda f, i: _RefGrouping.apply(f, i.int().contiguous(), self.R)  # the reference's own gather kernel
        try:
            return super().forward(xyz, mask, feats)
        finally:
            agg.gather = old

    __call__ = forward


def make_step(model, criterion, lr, weight_decay):
    """One reference training step on the GPU (train_dist.py:440-451); the caller switches the fused kernels off."""
    net = RefGpuUNet(model)
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)

    def step(pts, mask, feats, offs):
        opt.zero_grad(set_to_none=True)
        loss = criterion(net(pts, mask, feats).transpose(1, 2), offs, mask)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10)
        opt.step()
        return loss

    return step
