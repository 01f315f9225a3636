"""Weight-gradient GEMM of a 1x1 convolution on rows: dW = g^T x with R >> Cout, Cin.  Compares one cuBLAS GEMM with a
batched split over the row dimension (bmm over S row chunks + sum).  Diagnostic."""
import torch


def timeit(fn, iters=20):
    for _ in range(3):
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
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device("cuda:0")
    for R, co, ci in [(131072, 72, 72), (131072, 144, 72), (131072, 72, 3), (131072, 3, 72), (32768, 288, 144), (32768, 144, 144),
                      (8192, 576, 288), (2048, 1152, 576)]:
        g = torch.randn(R, co, device=dev)
        x = torch.randn(R, ci, device=dev)
        ref = g.t() @ x
        line = f"R={R:6d} {co:4d}x{ci:<4d} gemm {timeit(lambda: g.t() @ x):7.1f} us |"
        for S in (8, 16, 32, 64, 128, 256):
            if R % S:
                continue
            f = lambda: torch.bmm(g.view(S, R // S, co).transpose(1, 2), x.view(S, R // S, ci)).sum(0)
            err = ((f() - ref).abs().max() / ref.abs().max()).item()
            line += f" S{S}: {timeit(f):6.1f}"
        print(line, f"(rel err {err:.1e}; bytes {(g.numel() + x.numel()) * 4 / 1e6:.0f} MB)")


if __name__ == "__main__":
    main()
