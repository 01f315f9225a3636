"""GPU: fused BatchNorm1d (+ReLU, +residual) against torch.nn.BatchNorm1d / F.relu (the kernels the reference runs),
forward, backward, running statistics, eval mode.  fp32 tolerance: rtol 1e-5 / atol 1e-5 forward, rtol 1e-4 backward."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,N", [(4, 72, 2048), (2, 144, 513), (3, 1152, 64), (16, 72, 8192)])
@pytest.mark.parametrize("relu,residual", [(True, False), (False, False), (True, True)])
@pytest.mark.parametrize("channel_last", [False, True])
def test_fused_bn_matches_torch(cuda_device, B, C, N, relu, residual, channel_last):
    from deep3dpointclouddenoising_b200.fused import batch_norm_act
    torch.manual_seed(C + N)
    bn_ref = nn.BatchNorm1d(C, momentum=0.1).to(cuda_device)
    with torch.no_grad():
        bn_ref.weight.uniform_(0.5, 1.5)
        bn_ref.bias.uniform_(-0.5, 0.5)
    bn = nn.BatchNorm1d(C, momentum=0.1).to(cuda_device)
    bn.load_state_dict(bn_ref.state_dict())
    def make(scale, shift, grad):
        if channel_last:  # (B, C, N) view of (B, N, C) rows: the d3d_bn_act_cl_* kernels, no layout change
            t = (torch.randn(B, N, C, device=cuda_device) * scale + shift).requires_grad_(grad).permute(0, 2, 1)
            assert not t.is_contiguous()
            return t
        return (torch.randn(B, C, N, device=cuda_device) * scale + shift).requires_grad_(grad)

    x = make(2, 3, True)  # mean >> 0: exercises the shifted sums
    res = make(1, 0, True) if residual else None
    g = make(1, 0, False)
    for training in (True, False):
        bn.train(training)
        bn_ref.train(training)
        y_ref = bn_ref(x)
        if res is not None:
            y_ref = y_ref + res
        if relu:
            y_ref = torch.relu(y_ref)
        y = batch_norm_act(bn, x, relu=relu, residual=res)
        assert y.is_contiguous() != channel_last
        torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=2e-5)
        inputs = [x, bn.weight, bn.bias] + ([res] if res is not None else [])
        inputs_ref = [x, bn_ref.weight, bn_ref.bias] + ([res] if res is not None else [])
        grads = torch.autograd.grad(y, inputs, g)
        grads_ref = torch.autograd.grad(y_ref, inputs_ref, g)
        # an element whose pre-activation is within rounding of 0 may fall on either side of the ReLU kink in the two
        # implementations (1 element in 9.4 M at the largest size): such elements are excluded from the dx comparison
        for k, (a, b_) in enumerate(zip(grads, grads_ref)):
            if relu and a.shape == y_ref.shape:
                keep = (y_ref.detach().abs() > 1e-4) | ((y.detach() > 0) == (y_ref.detach() > 0))
                assert (~keep).float().mean().item() < 1e-5
                a, b_ = a * keep, b_ * keep
            # (a kink element also moves its channel's dgamma / dbeta by one term of the 131072-term sum: rtol 1e-3 there)
            rtol = 1e-3 if (relu and a.dim() == 1) else 1e-4
            torch.testing.assert_close(a, b_, rtol=rtol, atol=1e-4 * max(1.0, b_.abs().max().item()))
        torch.testing.assert_close(bn.running_mean, bn_ref.running_mean, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(bn.running_var, bn_ref.running_var, rtol=1e-5, atol=1e-6)
        assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked)


def test_model_with_and_without_fused_bn_agree(cuda_device):
    """Whole U-Net, fused BN kernels vs torch/cuDNN BatchNorm+ReLU.  TF32 is switched off for the convolutions so
    that only BN rounding differs; the deepest level still normalises over few samples (B*N/128), which amplifies
    1e-7 differences, hence the comparison in relative Frobenius norm."""
    from deep3dpointclouddenoising_b200 import synthetic
    from deep3dpointclouddenoising_b200.utils.config import runtime
    import bench
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        model, criterion, cfg = bench.build_model("pospool", 4096)
        model = model.to(cuda_device)
        batch = [torch.from_numpy(a).to(cuda_device) for a in synthetic.make_batch(3, 4, 4096, ragged=True)]
        outs = []
        state = {k: v.clone() for k, v in model.state_dict().items()}
        for flag in (True, False):
            runtime.fused_batchnorm = flag
            model.load_state_dict(state)
            model.zero_grad(set_to_none=True)
            pred = model(batch[0], batch[1], batch[2])
            loss = criterion(pred.transpose(1, 2), batch[3], batch[1])
            loss.backward()
            outs.append((pred.detach().clone(), loss.item(), {n: p.grad.clone() for n, p in model.named_parameters()}))
    finally:
        runtime.fused_batchnorm = True
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    rel = ((outs[0][0] - outs[1][0]).norm() / outs[1][0].norm()).item()
    assert rel < 1e-3, rel
    assert abs(outs[0][1] - outs[1][1]) < 1e-4 * max(1.0, abs(outs[1][1]))
    g0 = torch.cat([g.flatten() for g in outs[0][2].values()])
    g1 = torch.cat([g.flatten() for g in outs[1][2].values()])
    assert ((g0 - g1).norm() / g1.norm()).item() < 1e-2


def test_model_channel_last_matches_channel_major(cuda_device):
    """Whole U-Net with activations kept channel-last (row-major GEMM convolutions, d3d_bn_act_cl_*, aggregation
    kernels without transposition) against the channel-major path (Conv1d, d3d_bn_act_*, transposes around every
    aggregation).  fp32 everywhere (TF32 off); same amplification caveat as above."""
    from deep3dpointclouddenoising_b200 import synthetic
    from deep3dpointclouddenoising_b200.utils.config import runtime
    import bench
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        outs = []
        for operator in ("pospool", "pseudo_grid"):
            model, criterion, cfg = bench.build_model(operator, 4096)
            model = model.to(cuda_device)
            batch = [torch.from_numpy(a).to(cuda_device) for a in synthetic.make_batch(5, 4, 4096, ragged=True)]
            state = {k: v.clone() for k, v in model.state_dict().items()}
            for flag in (True, False):
                runtime.channel_last = flag
                model.load_state_dict(state)
                model.zero_grad(set_to_none=True)
                pred = model(batch[0], batch[1], batch[2])
                assert pred.shape == (4, 3, 4096)
                loss = criterion(pred.transpose(1, 2), batch[3], batch[1])
                loss.backward()
                outs.append((pred.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters()}))
            rel = ((outs[-2][0] - outs[-1][0]).norm() / outs[-1][0].norm()).item()
            assert rel < 1e-3, (operator, rel)
            num = sum(((outs[-2][1][n] - outs[-1][1][n]) ** 2).sum() for n in outs[-1][1]).sqrt().item()
            den = sum((outs[-1][1][n] ** 2).sum() for n in outs[-1][1]).sqrt().item()
            assert num / den < 2e-2, (operator, num / den)
    finally:
        runtime.channel_last = True
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("num_points", [2048, 4096])  # 4 x 4096 = 16384 rows: the split-K weight-gradient path
def test_in_place_parameter_gradients_match_autograd_accumulation(cuda_device, num_points):
    """runtime.grads_in_place: BN and 1x1-conv backward add the parameter gradients straight into the (flat) gradient
    buffers and return None to autograd — the buffers must equal what autograd's own accumulation produces.  The
    weight-gradient GEMMs run on a side stream then; FlatParameters.reduce() joins them."""
    from deep3dpointclouddenoising_b200 import distributed, synthetic
    from deep3dpointclouddenoising_b200.utils.config import runtime
    import bench
    model, criterion, cfg = bench.build_model("pospool", num_points)
    model = model.to(cuda_device)
    flat = distributed.FlatParameters(model)
    batch = [torch.from_numpy(a).to(cuda_device) for a in synthetic.make_batch(9, 4, num_points, ragged=True)]
    grads = []
    try:
        for flag in (False, True, True):  # twice in place: the second pass checks that the buffers really accumulate
            runtime.grads_in_place = flag
            if len(grads) < 2:
                flat.zero()
            loss = criterion(model(batch[0], batch[1], batch[2]).transpose(1, 2), batch[3], batch[1])
            loss.backward()
            flat.reduce()  # world size 1: only the join of the side-stream weight gradients
            grads.append(flat.flat.clone())
    finally:
        runtime.grads_in_place = False
    assert grads[0].abs().max() > 0
    torch.testing.assert_close(grads[1], grads[0], rtol=1e-4, atol=1e-5 * grads[0].abs().max().item())
    torch.testing.assert_close(grads[2], 2 * grads[0], rtol=1e-3, atol=1e-4 * grads[0].abs().max().item())
