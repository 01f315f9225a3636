"""CPU suite, part 5: the layout helpers and the row-major 1x1 convolutions of the channel-last pipeline
(fused.py: is_channel_last / rows_of / cat_channels; models/blocks.py: PointwiseConvRows, PointwiseConvCatRows, the
batched split-K weight gradient) against torch.nn.Conv1d / torch.cat — the modules they replace
(ref: u_net_arch/models/backbones/resnet.py:32-45, heads/multi_dimensional_head.py:36-37)."""
import pytest
import torch
import torch.nn as nn

from deep3dpointclouddenoising_b200.fused import cat_channels, is_channel_last, rows_of
from deep3dpointclouddenoising_b200.models import blocks


def _cl(b, c, n, grad=False):
    return torch.randn(b, n, c, dtype=torch.float64).requires_grad_(grad).permute(0, 2, 1)


def test_layout_helpers():
    x = _cl(2, 6, 5)
    assert is_channel_last(x) and not is_channel_last(x.contiguous())
    assert not is_channel_last(torch.randn(2, 6, 1).permute(0, 2, 1).permute(0, 2, 1))  # degenerate: treated as channel-major
    r = rows_of(x)
    assert r.is_contiguous() and r.data_ptr() == x.data_ptr() and r.shape == (2, 5, 6)
    r2 = rows_of(x.contiguous())
    assert r2.is_contiguous() and torch.equal(r2, r)
    y = _cl(2, 4, 5)
    z = cat_channels([x, y])
    assert is_channel_last(z) and torch.equal(z, torch.cat([x, y], 1))
    assert torch.equal(cat_channels([x.contiguous(), y]), torch.cat([x, y], 1))


def test_pointwise_conv_rows_matches_conv1d():
    torch.manual_seed(0)
    for bias in (False, True):
        conv = nn.Conv1d(6, 10, 1, bias=bias).double()
        x = _cl(3, 6, 7, grad=True)
        y = blocks.PointwiseConvRows.apply(rows_of(x), conv.weight, conv.bias).permute(0, 2, 1)
        ref = conv(x)
        torch.testing.assert_close(y, ref)
        g = torch.randn_like(ref)
        got = torch.autograd.grad(y, [x, conv.weight] + ([conv.bias] if bias else []), g)
        exp = torch.autograd.grad(ref, [x, conv.weight] + ([conv.bias] if bias else []), g)
        for a, b in zip(got, exp):
            torch.testing.assert_close(a, b)


def test_pointwise_conv_over_concatenation_matches_cat_then_conv1d():
    torch.manual_seed(1)
    conv = nn.Conv1d(9, 5, 1, bias=True).double()
    a, b = _cl(2, 4, 6, grad=True), _cl(2, 5, 6, grad=True)
    y = blocks.PointwiseConvCatRows.apply(conv.weight, conv.bias, False, a.permute(0, 2, 1), b.permute(0, 2, 1)).permute(0, 2, 1)
    ref = conv(torch.cat([a, b], 1))
    torch.testing.assert_close(y, ref)
    g = torch.randn_like(ref)
    got = torch.autograd.grad(y, [a, b, conv.weight, conv.bias], g)
    exp = torch.autograd.grad(ref, [a, b, conv.weight, conv.bias], g)
    for u, v in zip(got, exp):
        torch.testing.assert_close(u, v)


def test_split_weight_gradient_and_in_place_accumulation():
    torch.manual_seed(2)
    g, x = torch.randn(32768, 6, dtype=torch.float64), torch.randn(32768, 4, dtype=torch.float64)
    ref = g.t() @ x
    torch.testing.assert_close(blocks._weight_grad(g, x), ref)          # 32 row chunks + sum
    torch.testing.assert_close(blocks._weight_grad(g[:100], x[:100]), g[:100].t() @ x[:100])  # single GEMM
    buf = torch.ones(6, 4, dtype=torch.float64)
    assert blocks._weight_grad(g, x, into=buf) is None
    torch.testing.assert_close(buf, ref + 1)
    assert blocks._weight_grad(g[:100], x[:100], into=buf) is None
    torch.testing.assert_close(buf, ref + 1 + g[:100].t() @ x[:100])


def test_fused_sequential_refuses_cpu_unless_the_oracle_opts_in():
    """The product has no CPU path: the conv / BN blocks raise on CPU tensors.  With runtime.cpu_modules (set by the CPU
    oracle) a list input is the plain concatenation followed by the stock torch modules."""
    from deep3dpointclouddenoising_b200.utils.config import runtime
    torch.manual_seed(3)
    block = blocks.conv_bn(7, 4).double()
    a, b = torch.randn(2, 3, 9, dtype=torch.float64), torch.randn(2, 4, 9, dtype=torch.float64)
    block.train()
    with pytest.raises(RuntimeError, match="CPU not supported"):
        block([a, b])
    ref = nn.Sequential(*list(block.children()))(torch.cat([a, b], 1))
    runtime.cpu_modules = True
    try:
        torch.testing.assert_close(block([a, b]), ref)
    finally:
        runtime.cpu_modules = False
