"""Does spatial coherence of the point order speed up the gather kernels (L1 reuse between neighbouring queries)?
Times PosPool forward / backward at the level-0 shape on the synthetic batch as generated (random point order) and with
every cloud re-ordered along a Morton curve.  Diagnostic for the processing-order option of the aggregation kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import ops, synthetic  # noqa: E402


def morton_order(p, bits=10):
    q = ((p - p.min(0)) / (np.ptp(p, 0).max() + 1e-9) * ((1 << bits) - 1)).astype(np.uint64)
    code = np.zeros(len(p), dtype=np.uint64)
    for b in range(bits):
        for d in range(3):
            code |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + d)
    return np.argsort(code, kind="stable")


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
    dev = torch.device("cuda:0")
    B, N, C = 16, 8192, 72
    pts_np, mask_np, _, _ = synthetic.make_batch(1, B, N)
    for label in ("random order", "morton order"):
        if label == "morton order":
            pts_np = np.stack([p[morton_order(p)] for p in pts_np])
        pts, mask = torch.from_numpy(pts_np).to(dev), torch.from_numpy(mask_np).to(dev)
        idx, msk, nv = ops.ball_query(pts, pts, mask, mask, 0.025, 52, want_nvalid=True)
        rowptr, entries = ops.build_inverse_map(idx, N)
        for C in (72, 144):
            f = torch.randn(B, N, C, device=dev)
            t_f = timeit(lambda: ops.pospool_fwd(f, pts, pts, idx, nv, mask, 0.025, 'avg'))
            t_b = timeit(lambda: ops.pospool_bwd(f, pts, pts, rowptr, entries, nv, mask, N, 52, 0.025, 'avg'))
            print(f"{label}: C={C} pospool fwd {t_f:7.1f} us  bwd {t_b:7.1f} us   (mean nvalid {nv.float().mean().item():.1f})")
        t_q = timeit(lambda: ops.ball_query(pts, pts, mask, mask, 0.025, 52, want_nvalid=True), 5)
        print(f"{label}: ball query {t_q:7.1f} us")


if __name__ == "__main__":
    main()
