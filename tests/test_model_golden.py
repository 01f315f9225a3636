"""Whole-model parity against the REFERENCE'S OWN MODEL.

tests/golden/model_{pospool,pseudo_grid}.npz were written by oracle/make_golden_model.py: the reference's
u_net_arch/models (build.py:236-262, backbones/resnet.py:71-188, heads/multi_dimensional_head.py) imported from
/root/reference, one fp32 training step on the CPU at B=2, N=1024, ragged masks.  Here:

* CPU (not gpu): the oracle port of the step (oracle/cpu_model.py, the thing bench.py times as cpu_baseline)
  reproduces the reference's prediction, loss and every parameter gradient;
* GPU: the CUDA path (this package's model on the fused sm_100a kernels, TF32 off on both sides like the golden run)
  does too — fp32 tolerances for PosPool and fp32 PseudoGrid, the stated bf16 tolerance for the tcgen05 PseudoGrid.
"""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_model as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build(kind):
    from deep3dpointclouddenoising_b200.models import build_offset_regression
    c = G.make_config(kind)
    torch.manual_seed(0)
    model, criterion = build_offset_regression(c)
    missing = model.load_state_dict(G.seeded_state(model), strict=False)
    assert not missing.unexpected_keys
    return model.train(), criterion


def _compare(gold, pred, loss, named_grads, rtol_pred, rtol_grad, label, only=None):
    ref_pred = gold["pred"]
    scale = np.abs(ref_pred).max()
    err = np.abs(pred - ref_pred).max()
    assert err <= rtol_pred * scale, f"{label}: prediction differs by {err:.3e} (scale {scale:.3e})"
    assert abs(loss - float(gold["loss"])) <= rtol_pred * abs(float(gold["loss"])) * 4, (label, loss, float(gold["loss"]))
    checked, worst, errs = 0, 0.0, []
    for name, g in named_grads:
        checked += 1
        if only is not None and not name.startswith(only):
            continue
        flat = g.reshape(-1)
        want = gold["g:" + name]
        got = flat[G.sample_indices(name, flat.size)]
        norm_ref, sum_ref = gold["n:" + name]
        norm = float(np.sqrt((flat.astype(np.float64) ** 2).sum()))
        denom = max(np.abs(want).max(), norm_ref / np.sqrt(max(flat.size, 1)), 1e-12)
        e = np.abs(got - want).max() / denom
        errs.append(float(e))
        worst = max(worst, float(e))
        assert e <= rtol_grad, f"{label}: d/d{name} sampled entries differ by {e:.3e} of their scale"
        assert abs(norm - norm_ref) <= rtol_grad * max(norm_ref, 1e-12), f"{label}: |d/d{name}| {norm} vs {norm_ref}"
    assert checked == sum(1 for k in gold.files if k.startswith("g:")), "parameter sets differ"
    # the bulk of the parameters must sit an order of magnitude below the bound on the worst one (a wrong kernel moves
    # every gradient by O(1): see the bf16 note below)
    assert np.median(errs) <= rtol_grad / 10, f"{label}: median gradient error {np.median(errs):.3e}"
    print(f"{label}: gradient error median {np.median(errs):.2e} / worst {worst:.2e} of the tensor scale, "
          f"prediction error {err / scale:.2e}")


@pytest.mark.parametrize("kind", ["pospool", "pseudo_grid"])
def test_cpu_port_reproduces_reference_model_step(oracle, kind):
    from oracle.cpu_model import CpuUNet
    gold = np.load(os.path.join(GOLD, f"model_{kind}.npz"))
    model, criterion = _build(kind)
    pts, mask, feats, offs = [torch.from_numpy(a) for a in G.make_inputs()]
    pred = CpuUNet(model, oracle)(pts, mask, feats)
    loss = criterion(pred.transpose(1, 2), offs, mask)
    loss.backward()
    grads = [(n, p.grad.numpy()) for n, p in model.named_parameters()]
    _compare(gold, pred.detach().numpy(), loss.item(), grads, 2e-5, 2e-4, f"cpu port / {kind}")


@pytest.mark.gpu
@pytest.mark.parametrize("kind,precision,staged,rtol_pred,rtol_grad", [
    # Gradient tolerances follow the MEASURED fp32 sensitivity of the reference algorithm itself: evaluating the same
    # step in float64 (oracle port) moves the PosPool model's gradients by up to 1.2e-2 of a tensor's scale against the
    # fp32 reference golden (ReLU / max-pool kinks flip on 1e-7 perturbations of a mean over <= 52 neighbours), the
    # PseudoGrid model's by 4e-5 — so 1.2e-2 is the resolution of the PosPool golden.  The error of the WORST of the 109
    # tensors is heavy-tailed (a handful of kink flips at the deepest level, 32 rows per BatchNorm): measured here between
    # 3e-3 and 5.4e-2 depending on nothing but summation orders (BatchNorm backward reduction, cuBLAS algorithm choice,
    # float atomics), identical in distribution for the staged tiles and the gather kernels — so the worst tensor is held to
    # 1e-1 and the MEDIAN tensor to 1e-2 (measured 1.5e-3).  PseudoGrid fp32: 1e-3 / 1e-4 (measured 8e-5 / 4e-5).
    # Predictions: 1e-4 of the output scale (measured 4e-6).
    ("pospool", "fp32", True, 1e-4, 1e-1),
    ("pospool", "fp32", False, 1e-4, 1e-1),     # per-query gather kernels instead of the staged tiles
    ("pseudo_grid", "fp32", True, 1e-4, 1e-3),
    # tcgen05 contraction with bf16 operands, stated separately: 2e-2 per operator, forward and both gradients
    # (tests/test_gpu_aggregation.py; measured 2.5e-3).  Ten PseudoGrid layers in sequence are held to 5e-2 of the output
    # scale (measured 2.2e-2).  The same sensitivity that turns 1e-7 into 4e-5 above (x400, BatchNorm over the 32 rows
    # of the deepest level) turns the operator's 2.5e-3 into O(1) differences of the BACKBONE gradients on this tiny
    # geometry, so only the gradients of the head's last layers (measured 1.4e-2) are compared for bf16.
    ("pseudo_grid", "bf16", True, 5e-2, 1e-1),
])
def test_cuda_model_matches_reference_model_step(cuda_device, kind, precision, staged, rtol_pred, rtol_grad):
    from deep3dpointclouddenoising_b200.utils.config import runtime
    gold = np.load(os.path.join(GOLD, f"model_{kind}.npz"))
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, runtime.pseudo_grid_precision,
           runtime.staged_tiles)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    runtime.pseudo_grid_precision = precision
    runtime.staged_tiles = staged
    try:
        model, criterion = _build(kind)
        model = model.to(cuda_device)
        pts, mask, feats, offs = [torch.from_numpy(a).to(cuda_device) for a in G.make_inputs()]
        pred = model(pts, mask, feats)
        loss = criterion(pred.transpose(1, 2), offs, mask)
        loss.backward()
        torch.cuda.synchronize()
        grads = [(n, p.grad.cpu().numpy()) for n, p in model.named_parameters()]
        _compare(gold, pred.detach().cpu().numpy(), loss.item(), grads, rtol_pred, rtol_grad,
                 f"cuda / {kind} / {precision}", only="segmentation_head.head." if precision == "bf16" else None)
        sd = model.state_dict()
        for key in gold.files:  # BatchNorm running statistics after the step (momentum update of batch statistics)
            if key.startswith("s:"):
                got, want = sd[key[2:]].cpu().numpy(), gold[key]
                assert np.abs(got - want).max() <= max(rtol_pred, 1e-4) * max(np.abs(want).max(), 1e-6) * 10, key
    finally:
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, runtime.pseudo_grid_precision,
         runtime.staged_tiles) = old
