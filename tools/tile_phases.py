"""Phase timing of the staged-tile PosPool kernels at the level-0 shape: per-CTA %globaltimer stamps written by the
kernels themselves (d3d_pospool_tiles_debug_timing), averaged.  Diagnostic only.
usage: python tools/tile_phases.py [fwd|bwd]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import _lib, ops, synthetic  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dev = torch.device("cuda:0")
B, N, C = 16, 8192, 72
pts, mask, _, _ = [torch.from_numpy(a).to(dev) for a in synthetic.make_batch(1, B, N)]
idx, msk, nv, bys = ops.ball_query(pts, pts, mask, mask, 0.025, 52, want_nvalid=True, want_by_support=True)
order = ops.spatial_order(pts)
plan = ops.tile_plan(bys, nv, mask, order, N)
f = torch.randn(B, N, C, device=dev)
L = _lib.load()
n_cta = (N // 128) * B
buf = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)


def run():
    if which == "fwd":
        ops.pospool_fwd(f, pts, pts, idx, nv, mask, 0.025, 'avg', query_order=order, idx_by_support=bys, plan=plan)
    else:
        ops.pospool_bwd(f, pts, pts, None, None, nv, mask, N, 52, 0.025, 'avg', query_order=order, idx_by_support=bys, plan=plan,
                        ordered=False)


for _ in range(3):
    run()
torch.cuda.synchronize()
fn = L.d3d_pospool_tiles_debug_timing
fn.argtypes = [ctypes.c_void_p]
fn.restype = None
fn(buf.data_ptr())
run()
torch.cuda.synchronize()
fn(None)
t = buf.view(n_cta, 8).cpu().double()
names = (["owners / TMEM / barriers", "(unused)", "(unused)", "chunk loop", "epilogue"] if which == "fwd" else
         ["owners / TMEM / barriers", "wait for the gradient rows", "convert to planes + plan ranks", "A blocks / MMA / drain",
          "(end)"])
span = (t[:, 5].max() - t[:, 0].min()) / 1e3
print(f"{which}: {n_cta} CTAs, kernel span {span:.1f} us, union rows per tile mean {t[:, 6].mean():.0f} (max {t[:, 6].max():.0f})")
for k, nm in enumerate(names):
    d = (t[:, k + 1] - t[:, k]) / 1e3
    print(f"  {nm:28s} mean {d.mean():6.2f} us   p90 {d.quantile(0.9):6.2f}   max {d.max():6.2f}")
tot = (t[:, 5] - t[:, 0]) / 1e3
print(f"  {'CTA lifetime':28s} mean {tot.mean():6.2f} us   p90 {tot.quantile(0.9):6.2f}   max {tot.max():6.2f}")
chunks = torch.ceil(t[:, 6] / (32 if which == "fwd" else 128))
print(f"  chunk loop per chunk: {((t[:, 4] - t[:, 3]) / 1e3 / chunks.clamp_min(1)).mean():.2f} us")
