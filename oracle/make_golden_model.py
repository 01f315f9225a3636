"""TEST INFRASTRUCTURE — whole-model golden vectors from the REFERENCE'S OWN MODEL, run in the build container.

The reference's u_net_arch/models package (build.py:236-262 OffsetRegressionModel, backbones/resnet.py:71-188,
heads/multi_dimensional_head.py, local_aggregation_operators.py, pt_custom_ops/pt_utils.py) is imported from
/root/reference unmodified; `pt_custom_ops._ext` is a stub over the reference's CUDA kernel definitions compiled
for the host (oracle/_ref/libref_emul.so), `tkinter`, `easydict` and `pytorch3d` (absent from this image, only
needed by import lines off the path) are stub modules.  One training step — forward, MaskedL1Loss, backward —
runs on the CPU in fp32 for cfgs/l1.yaml (PseudoGrid) and cfgs/l1_pospool.yaml (PosPool) at B=2, N=1024 with
ragged masks, and writes tests/golden/model_<kind>.npz:

    pred (B, 3, N), loss, and for EVERY parameter its gradient's L2 norm, sum and 4096 seeded samples
    (the whole tensor when it has <= 4096 entries); running_mean / running_var of three BatchNorm layers.

Weights are not stored: `seeded_state` regenerates them from the parameter names (torch CPU generator), the
tests call the same function.  /root/reference does not exist on the GPU box, hence the committed vectors.

    make -C oracle all && python oracle/make_golden_model.py
"""
import hashlib
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OUT = os.path.join(ROOT, "tests", "golden")
SAMPLES = 4096
B, N = 2, 1024
NPOINTS = [256, 64, 32, 16]  # the l1.yaml ratios (N/4, N/16, N/32, N/64-ish) at N = 1024


def seeded_state(model):
    """Deterministic weights from the parameter NAMES (so both sides build the same state without a file):
    conv / kernel weights ~ N(0, 1/sqrt(fan_in)) , BN weight ~ 1 + 0.1 N(0,1), BN bias ~ 0.1 N(0,1)."""
    state = {}
    for name, t in model.state_dict().items():
        if not t.dtype.is_floating_point:
            continue
        if name.endswith("running_mean") or name.endswith("running_var") or name.endswith("K_points"):
            continue
        seed = int.from_bytes(hashlib.sha256(name.encode()).digest()[:4], "little")
        g = torch.Generator().manual_seed(seed)
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        if t.dim() >= 2:
            fan_in = t.shape[1] if t.dim() == 3 else t.shape[0]
            state[name] = r / float(np.sqrt(max(fan_in, 1)))
            if name.endswith("kernel_weights"):
                state[name] = r * 0.3
        elif name.endswith("weight"):
            state[name] = 1.0 + 0.1 * r
        else:
            state[name] = 0.1 * r
    return state


def sample_indices(name, numel):
    if numel <= SAMPLES:
        return np.arange(numel)
    seed = int.from_bytes(hashlib.sha256(("idx:" + name).encode()).digest()[:4], "little")
    return np.sort(np.random.default_rng(seed).choice(numel, SAMPLES, replace=False))


def make_config(kind):
    from deep3dpointclouddenoising_b200.utils import config as cfgmod
    cfgmod.reset_config()
    name = "l1.yaml" if kind == "pseudo_grid" else "l1_pospool.yaml"
    cfgmod.update_config(os.path.join(ROOT, "deep3dpointclouddenoising_b200", "cfgs", name))
    c = cfgmod.config
    c.num_points = N
    cfgmod.apply_train_geometry(c)
    c.npoints = list(NPOINTS)
    c.input_features_dim = 0
    return c


def make_inputs():
    from deep3dpointclouddenoising_b200 import synthetic
    return synthetic.make_batch(41, B, N, ragged=True)


def install_reference_models():
    from oracle import make_golden
    make_golden.install_reference_python()  # pt_custom_ops._ext stub, `models` namespace, easydict
    tk = types.ModuleType("tkinter")
    tk.OFF = 0
    sys.modules.setdefault("tkinter", tk)
    for mod in ("pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn", "pytorch3d.loss", "pytorch3d.structures",
                "pytorch3d.structures.pointclouds", "pytorch3d.loss.chamfer"):
        if mod not in sys.modules:
            m = types.ModuleType(mod)
            m.__path__ = []
            sys.modules[mod] = m
    sys.modules["pytorch3d.ops.knn"].knn_gather = None
    sys.modules["pytorch3d.ops.knn"].knn_points = None
    sys.modules["pytorch3d.ops"].knn_points = None
    sys.modules["pytorch3d.ops"].knn_gather = None
    sys.modules["pytorch3d.structures.pointclouds"].Pointclouds = type("Pointclouds", (), {})
    sys.modules["pytorch3d.structures"].Pointclouds = sys.modules["pytorch3d.structures.pointclouds"].Pointclouds
    return importlib.import_module("models.build")


def main():
    os.environ["JOB_LOAD_DIR"] = "/root/reference/u_net_arch"  # K_points fixtures of the reference
    build = install_reference_models()
    pts, mask, feats, offs = make_inputs()
    for kind in ("pospool", "pseudo_grid"):
        c = make_config(kind)
        torch.manual_seed(0)
        model, criterion = build.build_offset_regression(c)
        model.load_state_dict(seeded_state(model), strict=False)
        model.train()
        xyz, m, f, o = [torch.from_numpy(a) for a in (pts, mask, feats, offs)]
        pred = model(xyz, m, f)
        loss = criterion(pred.transpose(1, 2), o, m)
        loss.backward()
        out = {"pred": pred.detach().numpy(), "loss": np.float64(loss.item())}
        for name, p in model.named_parameters():
            g = p.grad.detach().reshape(-1).numpy()
            out["g:" + name] = g[sample_indices(name, g.size)]
            out["n:" + name] = np.array([np.sqrt((g.astype(np.float64) ** 2).sum()), g.astype(np.float64).sum()])
        sd = model.state_dict()
        for name in ("backbone.conv1.1", "backbone.la1.local_aggregation_operator.out_transform.0",
                     "segmentation_head.head.1"):
            for stat in ("running_mean", "running_var"):
                if f"{name}.{stat}" in sd:
                    out[f"s:{name}.{stat}"] = sd[f"{name}.{stat}"].numpy()
        path = os.path.join(OUT, f"model_{kind}.npz")
        np.savez_compressed(path, **out)
        print(path, f"loss {loss.item():.6f}", f"{len(list(model.named_parameters()))} parameters",
              f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
