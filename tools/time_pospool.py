"""Times PosPool forward / backward at the ten LocalAggregation shapes of the U-Net, staged-tile
tensor-core kernels against the per-query gather kernels, plus the Morton order and union statistics.
usage: python tools/time_pospool.py [--levels 0,1,..]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import ops, synthetic  # noqa: E402


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
    B, N0 = 16, 8192
    pts, mask, _, _ = [torch.from_numpy(a).to(dev) for a in synthetic.make_batch(1, B, N0)]
    levels = [(pts, mask)]
    dl, npoints = 0.05 / 32, [2048, 512, 256, 64]
    for m in npoints:
        dl *= 2
        levels.append(ops.grid_subsample(levels[-1][0], levels[-1][1], m, dl))
    nsamples, radius, width = [52, 39, 32, 26, 26], 0.025, 144
    shapes = [(0, 0, radius, nsamples[0], width // 2)]
    for s in range(4):
        shapes.append((s + 1, s, radius, nsamples[s], width // 2 * 2 ** (s + 1) // 1))
        radius *= 2
        shapes.append((s + 1, s + 1, radius, nsamples[s + 1], width // 2 * 2 ** (s + 1)))
    print("order kernel (level 0): %.1f us" % timeit(lambda: ops.spatial_order(pts)))
    for lq, ls, r, ns, C in shapes:
        (q, qm), (s, sm) = levels[lq], levels[ls]
        M, N = q.shape[1], s.shape[1]
        idx, msk, nv, bys = ops.ball_query(q, s, qm, sm, r, ns, want_nvalid=True, want_by_support=True)
        rowptr, entries = ops.build_inverse_map(idx, N)
        oq = ops.spatial_order(q)
        f = torch.randn(B, N, C, device=dev)
        g = torch.randn(B, M, C, device=dev)
        t = {}
        plan = ops.tile_plan(bys, nv, qm, oq, N)
        t["plan"] = timeit(lambda: ops.tile_plan(bys, nv, qm, oq, N))
        t["fwd gather"] = timeit(lambda: ops.pospool_fwd(f, q, s, idx, nv, qm, r, 'avg'))
        t["fwd tiles"] = timeit(lambda: ops.pospool_fwd(f, q, s, idx, nv, qm, r, 'avg', query_order=oq, idx_by_support=bys, plan=plan))
        t["bwd gather"] = timeit(lambda: ops.pospool_bwd(g, q, s, rowptr, entries, nv, qm, N, ns, r, 'avg'))
        t["bwd scatter"] = timeit(lambda: ops.pospool_bwd(g, q, s, None, None, nv, qm, N, ns, r, 'avg', query_order=oq, idx_by_support=bys, plan=plan, ordered=False))
        t["bwd ordered"] = timeit(lambda: ops.pospool_bwd(g, q, s, None, None, nv, qm, N, ns, r, 'avg', query_order=oq, idx_by_support=bys, plan=plan, ordered=True))
        a = ops.pospool_fwd(f, q, s, idx, nv, qm, r, 'avg')
        b = ops.pospool_fwd(f, q, s, idx, nv, qm, r, 'avg', query_order=oq, idx_by_support=bys)
        err = ((a - b).abs().max() / a.abs().max()).item()
        print(f"M={M:5d} N={N:5d} ns={ns} C={C:4d}: " + "  ".join(f"{k} {v:7.1f} us" for k, v in t.items()) + f"   max rel diff {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
