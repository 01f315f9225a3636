"""Times the fused BatchNorm kernels (channel-major vs channel-last entry points) per U-Net shape.  Diagnostic."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import ops  # noqa: E402


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    dev = torch.device("cuda:0")
    B = 16
    print(f"{'C':>5} {'N':>6} | {'cm fwd':>8} {'cl fwd':>8} | {'cm bwd':>8} {'cl bwd':>8}   (us)  MB/act")
    shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ["BN_SHAPES"].split(",")] if os.environ.get("BN_SHAPES") else None
    for C, N in shapes or [(72, 8192), (144, 8192), (144, 2048), (288, 2048), (288, 512), (576, 512), (576, 128), (1152, 128),
                 (1152, 32), (2304, 32)]:
        x = torch.randn(B, C, N, device=dev)
        xr = x.permute(0, 2, 1).contiguous()
        g = torch.randn(B, C, N, device=dev)
        gr = g.permute(0, 2, 1).contiguous()
        w, b_ = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        y, mean, invstd = ops.bn_act_fwd(x, None, w, b_, rm, rv, 1e-5, 0.1, True, True)
        t = [timeit(lambda: ops.bn_act_fwd(x, None, w, b_, rm, rv, 1e-5, 0.1, True, True)),
             timeit(lambda: ops.bn_act_cl_fwd(xr, None, w, b_, rm, rv, 1e-5, 0.1, True, True)),
             timeit(lambda: ops.bn_act_bwd(g, x, None, w, b_, mean, invstd, True, 1, False)),
             timeit(lambda: ops.bn_act_cl_bwd(gr, xr, None, w, b_, mean, invstd, True, 1, False))]
        print(f"{C:5d} {N:6d} | {t[0]:8.1f} {t[1]:8.1f} | {t[2]:8.1f} {t[3]:8.1f}          {x.numel() * 4 / 1e6:.1f}")


if __name__ == "__main__":
    main()
