"""How much do the neighbour lists of spatially adjacent queries overlap?  (evidence for the staged-tile
aggregation kernel planned next.)  Pure numpy on one synthetic 8192-point patch, reference selection rule restated:
first 3*ns in-radius supports in index order, then the ns nearest of those (masked_ordered_ball_query_gpu.cu:37-94).
usage: python tools/neighbor_overlap.py [group_size]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import synthetic  # noqa: E402


def main():
    group = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    pts, _, _, _ = synthetic.make_batch(1, 1, 8192)
    p = pts[0].astype(np.float64)
    radius, ns = 0.025, 52
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    inr = d2 < radius * radius
    cnt = inr.sum(1)
    print(f"supports inside the ball: mean {cnt.mean():.0f}, median {np.median(cnt):.0f}, min {cnt.min()}, max {cnt.max()}")
    nbrs, scanned = [], []
    for j in range(len(p)):
        c = np.nonzero(inr[j])[0]
        scanned.append((c[3 * ns - 1] + 1) / len(p) if len(c) >= 3 * ns else 1.0)
        c = c[:3 * ns]
        nbrs.append(c[np.argsort(d2[j, c], kind="stable")][:ns])
    print(f"fraction of the supports scanned until the candidate list is full: {np.mean(scanned):.2f}")
    rng = np.random.default_rng(0)
    unions = []
    for q in rng.integers(0, len(p), 40):
        members = np.argsort(d2[q])[:group]  # a query and its group-1 nearest queries: one spatial cell
        unions.append(len(set(np.concatenate([nbrs[j] for j in members]).tolist())))
    gathers = group * ns
    print(f"union of the neighbour rows of {group} adjacent queries: mean {np.mean(unions):.0f} "
          f"(min {min(unions)}, max {max(unions)}) for {gathers} gathers -> {gathers / np.mean(unions):.1f}x reuse; "
          f"{np.mean(unions) * 288 / 1024:.0f} KB at C = 72")


if __name__ == "__main__":
    main()
