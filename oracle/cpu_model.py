"""TEST INFRASTRUCTURE — the whole U-Net step on the CPU: the reference's call sequence
(u_net_arch/models/backbones/resnet.py:47-68,144-188, heads/multi_dimensional_head.py:62-85,
pt_custom_ops/pt_utils.py:122-148,192-238) restated with the CPU index ops (oracle/cpu_index_ops.py) and the
float oracle (oracle/aggregation_ref.py).  It walks the module tree of a model object handed to it (1x1
convolutions / BatchNorm are the same torch modules, on CPU) and recomputes every neighbour list per call,
like the reference does.

Used by tests (whole-model parity of the CUDA path), by bench.py's cpu_baseline leg and by
`bench.py --impl reference`.  Never imported by the product package.
"""
import numpy as np
import torch

from . import aggregation_ref as agg
from . import cpu_index_ops


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy())


class CpuUNet:
    def __init__(self, model, index_ops=None):
        self.model = model
        self.ix = index_ops if index_ops is not None else cpu_index_ops.restated()

    # --- index ops on torch CPU tensors
    def ball_query(self, q, s, qm, sm, radius, ns):
        idx, msk = self.ix.ball_query(_np(q), _np(s), _np(qm), _np(sm), radius, ns)
        return torch.from_numpy(idx), torch.from_numpy(msk)

    def local_aggregation(self, la, q_xyz, s_xyz, q_mask, s_mask, feats):
        op = la.local_aggregation_operator
        idx, msk = self.ball_query(q_xyz, s_xyz, q_mask, s_mask, op.radius, op.nsample)
        if hasattr(op, "kernel_weights"):
            out = agg.pseudogrid(feats, op.kernel_weights, op.K_points, q_xyz, s_xyz, q_mask, idx, msk, op.extent,
                                 op.KP_influence)
        else:
            out = agg.pospool(feats, q_xyz, s_xyz, q_mask, idx, msk, op.radius, op.reduction, op.position_embedding)
        return op.out_conv(out) if op.output_conv else op.out_transform(out)

    def max_pool(self, mp, xyz, mask, feats):
        sub, subm = self.ix.grid_subsampling(_np(xyz), _np(mask), mp.npoint, mp.sampleDl)
        sub, subm = torch.from_numpy(sub), torch.from_numpy(subm)
        idx, _ = self.ball_query(sub, xyz, subm, mask, mp.radius, mp.nsample)
        return sub, subm, agg.max_pool(feats, idx)

    def bottleneck(self, blk, xyz, mask, feats):
        if blk.downsample:
            q_xyz, q_mask, identity = self.max_pool(blk.maxpool, xyz, mask, feats)
        else:
            q_xyz, q_mask, identity = xyz, mask, feats
        out = blk.conv1(feats)
        out = self.local_aggregation(blk.local_aggregation, q_xyz, xyz, q_mask, mask, out)
        out = blk.conv2(out)
        if blk.in_channels != blk.out_channels:
            identity = blk.shortcut(identity)
        return q_xyz, q_mask, torch.relu(out + identity)

    def backbone(self, xyz, mask, feats):
        bb = self.model.backbone
        end = {}
        feats = bb.conv1(feats)
        feats = self.local_aggregation(bb.la1, xyz, xyz, mask, mask, feats)
        xyz, mask, feats = self.bottleneck(bb.btnk1, xyz, mask, feats)
        end["res1"] = (xyz, mask, feats)
        for stage in range(4):
            for blk in getattr(bb, f"layer{stage + 1}")._modules.values():
                xyz, mask, feats = self.bottleneck(blk, xyz, mask, feats)
            end[f"res{stage + 2}"] = (xyz, mask, feats)
        return end

    def head(self, end):
        hd = self.model.segmentation_head
        feats = end["res5"][2]
        for level in range(4):
            fine, coarse = end[f"res{4 - level}"], end[f"res{5 - level}"]
            nidx, _ = self.ix.nearest_query(_np(fine[0]), _np(coarse[0]), _np(fine[1]), _np(coarse[1]))
            feats = agg.nearest_upsample(feats, torch.from_numpy(nidx))
            feats = torch.cat([feats, fine[2]], 1)
            feats = getattr(hd, f"up_conv{level}")(feats)
        return hd.head(feats)

    def forward(self, xyz, mask, feats):
        # the module tree's 1x1 convolutions / BatchNorms are stock torch modules: the product refuses to run them on CPU
        # tensors unless the oracle says so (utils.config.runtime.cpu_modules)
        from deep3dpointclouddenoising_b200.utils.config import runtime
        old, runtime.cpu_modules = runtime.cpu_modules, True
        try:
            return self.head(self.backbone(xyz, mask, feats))
        finally:
            runtime.cpu_modules = old

    __call__ = forward

    def neighbour_build(self, xyz, mask, base_radius, base_dl, nsamples, npoints):
        """Only the index work of one forward (4 subsamplings, 9 distinct ball queries, 4 nearest queries)."""
        xyz, mask = _np(xyz), _np(mask)
        radius, dl = base_radius, base_dl
        self.ix.ball_query(xyz, xyz, mask, mask, radius, nsamples[0])
        levels = [(xyz, mask)]
        for stage in range(4):
            dl *= 2
            sub, subm = self.ix.grid_subsampling(xyz, mask, npoints[stage], dl)
            self.ix.ball_query(sub, xyz, subm, mask, radius, nsamples[stage])
            radius *= 2
            self.ix.ball_query(sub, sub, subm, subm, radius, nsamples[stage + 1])
            xyz, mask = sub, subm
            levels.append((xyz, mask))
        for fine, coarse in zip(levels[-2::-1], levels[:0:-1]):
            self.ix.nearest_query(fine[0], coarse[0], fine[1], coarse[1])
        return levels
