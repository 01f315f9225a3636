"""GPU: the whole U-Net through the public model API — forward/backward run, neighbour cache on/off gives
identical results, gradients reach every parameter, one optimiser step changes the loss."""
import numpy as np
import pytest
import torch

from deep3dpointclouddenoising_b200 import synthetic

pytestmark = pytest.mark.gpu


def _build(kind, num_points, device):
    from deep3dpointclouddenoising_b200.utils import config as cfgmod
    from deep3dpointclouddenoising_b200.models import build_offset_regression
    import os
    cfgmod.reset_config()
    name = "l1.yaml" if kind == "pseudo_grid" else "l1_pospool.yaml"
    cfgmod.update_config(os.path.join(os.path.dirname(cfgmod.__file__), "..", "cfgs", name))
    c = cfgmod.config
    c.num_points = num_points
    cfgmod.apply_train_geometry(c)
    c.input_features_dim = 0
    torch.manual_seed(0)
    np.random.seed(0)
    model, criterion = build_offset_regression(c)
    model.init_weights()
    return model.to(device), criterion


@pytest.mark.parametrize("kind", ["pospool", "pseudo_grid"])
def test_forward_backward_and_cache_equivalence(cuda_device, kind):
    from deep3dpointclouddenoising_b200 import neighbors
    from deep3dpointclouddenoising_b200.utils.config import runtime
    B, N = 2, 1024
    model, criterion = _build(kind, N, cuda_device)
    pts, mask, feats, offs = [torch.from_numpy(a).to(cuda_device) for a in synthetic.make_batch(5, B, N, ragged=True)]
    results = []
    # bit-for-bit comparison: PosPool backward without float atomics (the default scatter tiles agree to rounding only)
    old, runtime.staged_tiles_backward = runtime.staged_tiles_backward, False
    try:
        for cache_on in (True, False):
            neighbors.cache.enabled = cache_on
            model.zero_grad(set_to_none=True)
            pred = model(pts, mask, feats)
            assert pred.shape == (B, 3, N) and torch.isfinite(pred).all()
            loss = criterion(pred.transpose(1, 2), offs, mask)
            loss.backward()
            grads = {n: p.grad.clone() for n, p in model.named_parameters()}
            assert all(g is not None and torch.isfinite(g).all() for g in grads.values())
            results.append((pred.detach().clone(), loss.item(), grads))
    finally:
        neighbors.cache.enabled = True
        runtime.staged_tiles_backward = old
    assert torch.equal(results[0][0], results[1][0]) and results[0][1] == results[1][1]
    for n in results[0][2]:
        assert torch.equal(results[0][2][n], results[1][2][n]), n
    assert neighbors.cache.hits > 0


def test_one_adam_step_reduces_l1_on_fixed_batch(cuda_device):
    model, criterion = _build("pospool", 1024, cuda_device)
    pts, mask, feats, offs = [torch.from_numpy(a).to(cuda_device) for a in synthetic.make_batch(6, 4, 1024)]
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss = criterion(model(pts, mask, feats).transpose(1, 2), offs, mask)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10)
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


def test_deterministic_mode_gives_identical_training_steps(cuda_device):
    """utils.config.set_deterministic(): two runs of the same training step produce the same bits (loss and every gradient)."""
    from deep3dpointclouddenoising_b200.utils.config import runtime, set_deterministic
    old = (runtime.staged_tiles_backward, runtime.deterministic_scatter)
    set_deterministic(True)
    try:
        model, criterion = _build("pospool", 2048, cuda_device)
        pts, mask, feats, offs = [torch.from_numpy(a).to(cuda_device) for a in synthetic.make_batch(8, 2, 2048, ragged=True)]
        runs = []
        for _ in range(2):
            model.zero_grad(set_to_none=True)
            loss = criterion(model(pts, mask, feats).transpose(1, 2), offs, mask)
            loss.backward()
            runs.append((loss.item(), [p.grad.clone() for p in model.parameters()]))
        assert runs[0][0] == runs[1][0]
        assert all(torch.equal(a, b) for a, b in zip(runs[0][1], runs[1][1]))
    finally:
        runtime.staged_tiles_backward, runtime.deterministic_scatter = old
