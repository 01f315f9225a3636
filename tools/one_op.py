"""Runs ONE op of the hot path a few times at the BASELINE level-0 shape (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import ops, synthetic
op = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
C = int(sys.argv[3]) if len(sys.argv) > 3 else 72
dev = torch.device("cuda:0")
B, N = 16, 8192
pts, mask, feats, offs = [torch.from_numpy(a).to(dev) for a in synthetic.make_batch(1, B, N)]
idx, msk, nv, bys = ops.ball_query(pts, pts, mask, mask, 0.025, 52, want_nvalid=True, want_by_support=True)
f = torch.randn(B, N, C, device=dev)
kp = torch.randn(15, 3, device=dev) * 0.006; w = torch.randn(15, C, device=dev)
rowptr, entries = ops.build_inverse_map(idx, N)
order = ops.spatial_order(pts)
wg = torch.randn(144, C, device=dev)
plan = ops.tile_plan(bys, nv, mask, order, N)
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
ybn, mean, invstd = ops.bn_act_cl_fwd(f, None, gam, bet, rm, rv, 1e-5, 0.1, True, True)
gbn = torch.randn_like(f)
torch.cuda.synchronize()
for _ in range(n):
    if op == "ball_query": ops.ball_query(pts, pts, mask, mask, 0.025, 52)
    elif op == "inverse_map": ops.build_inverse_map(idx, N)
    elif op == "pospool_fwd": ops.pospool_fwd(f, pts, pts, idx, nv, mask, 0.025, 'avg')
    elif op == "pospool_bwd": ops.pospool_bwd(f, pts, pts, rowptr, entries, nv, mask, N, 52, 0.025, 'avg')
    elif op == "pospool_tiles_fwd": ops.pospool_fwd(f, pts, pts, idx, nv, mask, 0.025, 'avg', query_order=order, idx_by_support=bys, plan=plan)
    elif op == "pospool_scatter_bwd": ops.pospool_bwd(f, pts, pts, None, None, nv, mask, N, 52, 0.025, 'avg', query_order=order, idx_by_support=bys, plan=plan)
    elif op == "tile_plan": ops.tile_plan(bys, nv, mask, order, N)
    elif op == "bn_bwd": ops.bn_act_cl_bwd(gbn, f, ybn, gam, bet, mean, invstd, True, 1, False)
    elif op == "bn_fwd": ops.bn_act_cl_fwd(f, None, gam, bet, rm, rv, 1e-5, 0.1, True, True)
    elif op == "gemm": ops.gemm_tf32(f.view(-1, C), wg)
    elif op == "gemm_stats": ops.gemm_tf32(f.view(-1, C), wg, want_stats=True)
    elif op == "pseudogrid_fwd": ops.pseudogrid_fwd(f, pts, pts, idx, nv, mask, kp, w, 0.01, 'linear', 0)
    elif op == "pseudogrid_fwd_tc": ops.pseudogrid_fwd(f, pts, pts, idx, nv, mask, kp, w, 0.01, 'linear', 1)
    elif op == "pseudogrid_bwd": ops.pseudogrid_bwd(f, f, pts, pts, idx, rowptr, entries, nv, mask, kp, w, 0.01, 'linear', 0)
    elif op == "pseudogrid_bwd_tc": ops.pseudogrid_bwd(f, f, pts, pts, idx, rowptr, entries, nv, mask, kp, w, 0.01, 'linear', 1)
torch.cuda.synchronize()
print("ok", op)
