"""Times the TF32 tensor-core GEMM (csrc/gemm.cu) against torch / cuBLAS TF32 at the convolution shapes of the U-Net
(B = 16 x 8192 points, width 144).   usage: python tools/time_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import ops  # noqa: E402


def timeit(fn, iters=20):
    """Device time per call: the calls are captured in a CUDA graph, so host launch overhead is not in the number."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = True
    shapes = [(131072, 72, 72), (131072, 72, 144), (131072, 144, 144), (32768, 144, 288), (32768, 288, 144), (8192, 288, 576),
              (8192, 576, 288), (4096, 576, 1152), (4096, 1152, 576), (1024, 1152, 2304), (1024, 2304, 1152), (4096, 3456, 576),
              (131072, 216, 72)]
    tot_o = tot_t = 0.0
    for M, K, N in shapes:
        a = torch.randn(M, K, device=dev)
        b = torch.randn(N, K, device=dev)
        t_o = timeit(lambda: ops.gemm_tf32(a, b))
        t_s = timeit(lambda: ops.gemm_tf32(a, b, want_stats=True))
        t_t = timeit(lambda: torch.nn.functional.linear(a, b))
        fl = 2.0 * M * K * N
        by = 4.0 * (M * K + N * K + M * N)
        print(f"M={M:6d} K={K:4d} N={N:4d}: ours {t_o:7.1f} us ({fl / t_o / 1e6:6.1f} TF/s, {by / t_o / 1e3:6.0f} GB/s)  "
              f"+stats {t_s:7.1f} us  cuBLAS tf32 {t_t:7.1f} us", flush=True)
        tot_o += t_o
        tot_t += t_t
    print(f"sum: ours {tot_o:.1f} us, cuBLAS {tot_t:.1f} us")
    print("weight gradients dW = dY^T X (ours: tensor-core split-K + reduce; torch: the batched split-K bmm + sum of models/blocks.py)")
    from deep3dpointclouddenoising_b200.models import blocks
    from deep3dpointclouddenoising_b200.utils.config import runtime
    runtime.own_wgrad = False  # the torch leg below must be torch's
    tot_o = tot_t = 0.0
    for M, K, N in shapes:
        dy = torch.randn(M, N, device=dev)
        x = torch.randn(M, K, device=dev)
        t_o = timeit(lambda: ops.wgrad_tf32(dy, x))
        t_t = timeit(lambda: blocks._weight_grad(dy, x))
        print(f"R={M:6d} Cin={K:4d} Cout={N:4d}: ours {t_o:7.1f} us ({4.0 * M * (K + N) / t_o / 1e3:6.0f} GB/s)  torch {t_t:7.1f} us", flush=True)
        tot_o += t_o
        tot_t += t_t
    print(f"sum: ours {tot_o:.1f} us, torch {tot_t:.1f} us")


if __name__ == "__main__":
    main()
