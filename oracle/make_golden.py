"""TEST INFRASTRUCTURE — generates tests/golden/*.npz from the REFERENCE ITSELF, run in the build container.

  index ops     : the reference's CUDA kernel definitions compiled for the host (oracle/_ref/libref_emul.so,
                  built by oracle/Makefile from /root/reference sources where they lie)
  aggregation   : the reference's own Python modules, imported from /root/reference
                  (u_net_arch/pt_custom_ops/pt_utils.py, u_net_arch/models/local_aggregation_operators.py),
                  with `pt_custom_ops._ext` provided by a stub that calls the same host-compiled reference
                  kernels (the real extension refuses CPU tensors, group_points.cpp:36); outputs and
                  autograd gradients of PosPool / PseudoGrid / MaskedMaxPool / MaskedUpsample.

/root/reference does not exist on the GPU box, so the vectors are committed; this script is how they were
made:   make -C oracle all && python oracle/make_golden.py
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu_index_ops  # noqa: E402
from deep3dpointclouddenoising_b200 import synthetic  # noqa: E402
from deep3dpointclouddenoising_b200.utils.config import AttrDict  # noqa: E402

REF = "/root/reference/u_net_arch"
OUT = os.path.join(ROOT, "tests", "golden")


def install_reference_python():
    """Makes `import models.local_aggregation_operators` resolve to the reference's files."""
    E = cpu_index_ops.reference()

    def t(a, like):
        return torch.from_numpy(np.ascontiguousarray(a)).to(like.device)

    ext = types.ModuleType("pt_custom_ops._ext")
    ext.group_points = lambda p, i: t(E.group_points(p.detach().numpy(), i.numpy()), p)
    ext.group_points_grad = lambda g, i, n: t(E.group_points_grad(g.detach().numpy(), i.numpy(), n), g)
    ext.masked_ordered_ball_query = lambda q, s, qm, sm, r, ns: [
        t(a, q) for a in E.ball_query(q.numpy(), s.numpy(), qm.numpy(), sm.numpy(), r, ns)]
    ext.masked_nearest_query = lambda q, s, qm, sm: [t(a, q) for a in E.nearest_query(q.numpy(), s.numpy(), qm.numpy(), sm.numpy())]
    ext.masked_grid_subsampling = lambda p, m, n, dl: [t(a, p) for a in E.grid_subsampling(p.numpy(), m.numpy(), n, dl)]
    pkg = types.ModuleType("pt_custom_ops")
    pkg.__path__ = [os.path.join(REF, "pt_custom_ops")]
    pkg._ext = ext
    sys.modules["pt_custom_ops"], sys.modules["pt_custom_ops._ext"] = pkg, ext
    sys.path.insert(0, os.path.join(REF, "pt_custom_ops"))  # `from pt_utils import ...`
    models = types.ModuleType("models")  # skip models/__init__.py: it pulls pytorch3d / tkinter
    models.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = models
    easydict = types.ModuleType("easydict")
    easydict.EasyDict = AttrDict
    sys.modules["easydict"] = easydict
    sys.path.insert(0, REF)
    pt_utils = importlib.import_module("pt_utils")
    lao = importlib.import_module("models.local_aggregation_operators")
    return pt_utils, lao


def seeded_levels(seed, B, N, ragged=True):
    pts, mask, feats, _ = synthetic.make_batch(seed, B, N, ragged=ragged)
    return pts, mask


def index_goldens():
    E = cpu_index_ops.reference()
    cases = {}
    # (name, B, N, M or None for self query, radius, nsample)
    pts, mask = seeded_levels(11, 3, 1024)
    for name, radius, ns in (("self_r025_ns52", 0.025, 52), ("self_r05_ns16", 0.05, 16), ("self_r005_ns8", 0.005, 8)):
        idx, msk = E.ball_query(pts, pts, mask, mask, radius, ns)
        cases[f"bq_{name}_idx"], cases[f"bq_{name}_mask"] = idx, msk
    sub, subm = E.grid_subsampling(pts, mask, 256, 0.003125)
    cases["gs_dl003125_m256_xyz"], cases["gs_dl003125_m256_mask"] = sub, subm
    sub2, subm2 = E.grid_subsampling(pts, mask, 1024, 0.0125)  # fewer cells than m: cyclic padding
    cases["gs_dl0125_m1024_xyz"], cases["gs_dl0125_m1024_mask"] = sub2, subm2
    idx, msk = E.ball_query(sub, pts, subm, mask, 0.025, 52)
    cases["bq_sub_r025_ns52_idx"], cases["bq_sub_r025_ns52_mask"] = idx, msk
    nidx, nmsk = E.nearest_query(pts, sub, mask, subm)
    cases["nn_idx"], cases["nn_mask"] = nidx, nmsk
    cases["points"], cases["mask"] = pts, mask
    np.savez_compressed(os.path.join(OUT, "index_ops.npz"), **cases)
    print("index_ops.npz", {k: v.shape for k, v in cases.items()})


def aggregation_goldens():
    pt_utils, lao = install_reference_python()
    torch.manual_seed(0)
    np.random.seed(0)
    B, N, C, ns, radius = 2, 384, 24, 20, 0.025
    pts, mask = seeded_levels(23, B, N)
    xyz, m = torch.from_numpy(pts), torch.from_numpy(mask)
    sub_xyz, sub_mask = [torch.from_numpy(a) for a in cpu_index_ops.reference().grid_subsampling(pts, mask, 96, 0.00625)]
    out = {"points": pts, "mask": mask, "sub_xyz": sub_xyz.numpy(), "sub_mask": sub_mask.numpy()}
    feats = torch.randn(B, C, N)
    out["features"] = feats.numpy()

    def run(module, args, feat, params=()):
        feat = feat.clone().requires_grad_(True)
        y = module(*args, feat)
        g = torch.randn_like(y)
        grads = torch.autograd.grad(y, (feat,) + tuple(params), g)
        return y.detach().numpy(), g.numpy(), [x.numpy() for x in grads]

    cfg = AttrDict(bn_momentum=0.1, density_parameter=5.0,
                   pospool=AttrDict(position_embedding='xyz', reduction='avg', output_conv=False),
                   pseudo_grid=AttrDict(fixed_kernel_points='center', KP_influence='linear', KP_extent=1.0,
                                        num_kernel_points=15, convolution_mode='sum', output_conv=False))
    # the operators alone (BN/ReLU blocks replaced by identity so the goldens pin the aggregation itself)
    for tag, q, qm in (("self", xyz, m), ("strided", sub_xyz, sub_mask)):
        for red in ("avg", "sum", "max"):
            cfg.pospool.reduction = red
            op = lao.PosPool(C, C, radius, ns, cfg)
            op.out_transform = torch.nn.Identity()
            y, g, (gf,) = run(op, (q, xyz, qm, m), feats)
            out[f"pospool_{tag}_{red}_out"], out[f"pospool_{tag}_{red}_gout"], out[f"pospool_{tag}_{red}_gfeat"] = y, g, gf
        cfg.pospool.reduction, cfg.pospool.position_embedding = "avg", "sin_cos"
        op = lao.PosPool(C, C, radius, ns, cfg)
        op.out_transform = torch.nn.Identity()
        y, g, (gf,) = run(op, (q, xyz, qm, m), feats)
        out[f"pospool_{tag}_sincos_out"], out[f"pospool_{tag}_sincos_gout"], out[f"pospool_{tag}_sincos_gfeat"] = y, g, gf
        cfg.pospool.position_embedding = "xyz"
        for infl in ("linear", "constant"):  # "gaussian" raises TypeError in the reference itself (utlis.py:294 torch.pow(float, int))
            cfg.pseudo_grid.KP_influence = infl
            os.environ["JOB_LOAD_DIR"] = REF  # K_points from the reference's own fixture folder
            op = lao.PseudoGrid(C, C, radius, ns, cfg)
            op.out_transform = torch.nn.Identity()
            y, g, (gf, gw) = run(op, (q, xyz, qm, m), feats, (op.kernel_weights,))
            pre = f"pseudogrid_{tag}_{infl}"
            out[pre + "_out"], out[pre + "_gout"], out[pre + "_gfeat"], out[pre + "_gweights"] = y, g, gf, gw
            out[pre + "_weights"], out[pre + "_kpoints"] = op.kernel_weights.detach().numpy(), op.K_points.numpy()
            out[pre + "_extent"] = np.float32(op.extent)
    # MaskedMaxPool (its own subsampling + ball query) and nearest MaskedUpsample
    pool = pt_utils.MaskedMaxPool(96, radius, ns, 0.00625)
    f = feats.clone().requires_grad_(True)
    sx, sm_, sf = pool(xyz, m, f)
    g = torch.randn_like(sf)
    out["maxpool_sub_xyz"], out["maxpool_sub_mask"], out["maxpool_out"] = sx.numpy(), sm_.numpy(), sf.detach().numpy()
    out["maxpool_gout"], out["maxpool_gfeat"] = g.numpy(), torch.autograd.grad(sf, f, g)[0].numpy()
    up = pt_utils.MaskedUpsample(radius, ns, mode='nearest')
    coarse = torch.randn(B, C, 96)
    cf = coarse.clone().requires_grad_(True)
    y = up(xyz, sub_xyz, m, sub_mask, cf)
    g = torch.randn_like(y)
    out["upsample_features"], out["upsample_out"], out["upsample_gout"] = coarse.numpy(), y.detach().numpy(), g.numpy()
    out["upsample_gfeat"] = torch.autograd.grad(y, cf, g)[0].numpy()
    out["meta"] = np.array([B, N, C, ns], np.int64)
    out["radius"] = np.float32(radius)
    np.savez_compressed(os.path.join(OUT, "aggregation.npz"), **out)
    print("aggregation.npz", len(out), "arrays")


def extra_goldens():
    """Variants no head of the reference uses but its API offers (pt_utils.py:158-180, 227-234): MaskedUpsample
    'max' / 'rbf' and MaskedNearestQueryAndGroup.forward, outputs and autograd gradients from the reference's Python."""
    pt_utils, _ = install_reference_python()
    torch.manual_seed(1)
    # radius 0.06: every fine point has a coarse support inside the ball (with none the reference kernel computes
    # `i % 0`, masked_ordered_ball_query_gpu.cu:83-86 — undefined on the GPU, SIGFPE in the host build)
    B, N, C, ns, radius = 2, 384, 24, 20, 0.06
    pts, mask = seeded_levels(29, B, N)
    xyz, m = torch.from_numpy(pts), torch.from_numpy(mask)
    sub_xyz, sub_mask = [torch.from_numpy(a) for a in cpu_index_ops.reference().grid_subsampling(pts, mask, 96, 0.00625)]
    out = {"points": pts, "mask": mask, "sub_xyz": sub_xyz.numpy(), "sub_mask": sub_mask.numpy(),
           "meta": np.array([B, N, C, ns], np.int64), "radius": np.float32(radius)}
    coarse = torch.randn(B, C, 96)
    out["coarse"] = coarse.numpy()
    for mode in ("max", "rbf"):
        up = pt_utils.MaskedUpsample(radius, ns, mode=mode)
        cf = coarse.clone().requires_grad_(True)
        y = up(xyz, sub_xyz, m, sub_mask, cf)
        g = torch.randn_like(y)
        out[f"up_{mode}_out"], out[f"up_{mode}_gout"] = y.detach().numpy(), g.numpy()
        out[f"up_{mode}_gfeat"] = torch.autograd.grad(y, cf, g)[0].numpy()
    for use_xyz in (True, False):
        grouper = pt_utils.MaskedNearestQueryAndGroup(use_xyz=use_xyz, ret_grouped_xyz=True)
        cf = coarse.clone().requires_grad_(True)
        feats, gxyz, imask = grouper(xyz, sub_xyz, m, sub_mask, cf)
        g = torch.randn_like(feats)
        tag = "xyz" if use_xyz else "noxyz"
        out[f"nqg_{tag}_feat"], out[f"nqg_{tag}_gxyz"], out[f"nqg_{tag}_mask"] = feats.detach().numpy(), gxyz.numpy(), imask.numpy()
        out[f"nqg_{tag}_gout"], out[f"nqg_{tag}_gfeat"] = g.numpy(), torch.autograd.grad(feats, cf, g)[0].numpy()
    grouper = pt_utils.MaskedNearestQueryAndGroup(use_xyz=True)
    feats, imask = grouper(xyz, sub_xyz, m, sub_mask, None)
    out["nqg_nofeat_feat"] = feats.numpy()
    np.savez_compressed(os.path.join(OUT, "aggregation_extra.npz"), **out)
    print("aggregation_extra.npz", len(out), "arrays")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--extra-only" not in sys.argv:
        index_goldens()
        aggregation_goldens()
    extra_goldens()
