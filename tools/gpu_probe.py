"""Ad-hoc GPU probe: per-kernel CUDA-event timings at the BASELINE level-0 shapes + one model step."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import ops, synthetic, neighbors, fused

dev = torch.device("cuda:0")
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

B, N = 16, 8192
pts, mask, feats, offs = [torch.from_numpy(a).to(dev) for a in synthetic.make_batch(1, B, N)]
print("ball_query 8192x8192 ns52: %.3f ms" % timeit(lambda: ops.ball_query(pts, pts, mask, mask, 0.025, 52)))
sub, subm = ops.grid_subsample(pts, mask, 2048, 0.003125)
print("grid_subsample 8192->2048: %.3f ms" % timeit(lambda: ops.grid_subsample(pts, mask, 2048, 0.003125)))
print("ball_query 2048x8192 ns52: %.3f ms" % timeit(lambda: ops.ball_query(sub, pts, subm, mask, 0.025, 52)))
print("nearest 8192x2048: %.3f ms" % timeit(lambda: ops.nearest_query(pts, sub, mask, subm)))
idx, msk, nv = ops.ball_query(pts, pts, mask, mask, 0.025, 52, want_nvalid=True)
print("inverse map: %.3f ms" % timeit(lambda: ops.build_inverse_map(idx, N)))
rowptr, entries = ops.build_inverse_map(idx, N)
seg = (rowptr[1:] - rowptr[:-1]).float()
print("segment len: mean %.1f max %d  >64: %d  >1024: %d" % (seg.mean().item(), int(seg.max().item()), int((seg > 64).sum()), int((seg > 1024).sum())))
for C in (72, 144):
    f = torch.randn(B, C, N, device=dev)
    fcl = ops.cm_to_cl(f)
    print("C=%d cm_to_cl: %.3f ms" % (C, timeit(lambda: ops.cm_to_cl(f))))
    t = timeit(lambda: ops.pospool_fwd(fcl, pts, pts, idx, nv, mask, 0.025, 'avg'))
    print("C=%d pospool_fwd: %.3f ms  (gather %.1f GB -> %.0f GB/s L2-level)" % (C, t, B*N*52*C*4/1e9, B*N*52*C*4/1e6/t))
    g = torch.randn(B, N, C, device=dev)
    t = timeit(lambda: ops.pospool_bwd(g, pts, pts, rowptr, entries, nv, mask, N, 52, 0.025, 'avg'))
    print("C=%d pospool_bwd: %.3f ms" % (C, t))
    kp = torch.randn(15, 3, device=dev) * 0.01; w = torch.randn(15, C, device=dev)
    t = timeit(lambda: ops.pseudogrid_fwd(fcl, pts, pts, idx, nv, mask, kp, w, 0.01, 'linear', 0))
    print("C=%d pseudogrid_fwd fp32: %.3f ms" % (C, t))
    t = timeit(lambda: ops.pseudogrid_fwd(fcl, pts, pts, idx, nv, mask, kp, w, 0.01, 'linear', 1))
    print("C=%d pseudogrid_fwd tcgen05: %.3f ms" % (C, t))
    t = timeit(lambda: ops.pseudogrid_bwd(g, fcl, pts, pts, idx, rowptr, entries, nv, mask, kp, w, 0.01, 'linear', 0, True, False))
    print("C=%d pseudogrid_bwd feat fp32: %.3f ms" % (C, t))
    t = timeit(lambda: ops.pseudogrid_bwd(g, fcl, pts, pts, idx, rowptr, entries, nv, mask, kp, w, 0.01, 'linear', 1, True, False))
    print("C=%d pseudogrid_bwd feat tcgen05: %.3f ms" % (C, t))
    t = timeit(lambda: ops.pseudogrid_bwd(g, fcl, pts, pts, idx, rowptr, entries, nv, mask, kp, w, 0.01, 'linear', 0, False, True))
    print("C=%d pseudogrid_bwd weights: %.3f ms" % (C, t))
    t = timeit(lambda: ops.pseudogrid_bwd(g, fcl, pts, pts, idx, rowptr, entries, nv, mask, kp, w, 0.01, 'linear', 1, False, True))
    print("C=%d pseudogrid_bwd weights tcgen05: %.3f ms" % (C, t))
    t = timeit(lambda: ops.gather_max_fwd(fcl, idx))
    print("C=%d gather_max_fwd: %.3f ms" % (C, t))
f = torch.randn(B, 72, N, device=dev)
print("group_points C=72: %.3f ms" % timeit(lambda: ops.group_points(f, idx), n=3, warm=1))

# model step
from deep3dpointclouddenoising_b200.utils import config as cfgmod
from deep3dpointclouddenoising_b200.models import build_offset_regression
for name in ("l1_pospool.yaml", "l1.yaml"):
    cfgmod.reset_config(); cfgmod.update_config(os.path.join(os.path.dirname(cfgmod.__file__), "..", "cfgs", name))
    c = cfgmod.config; c.num_points = N; cfgmod.apply_train_geometry(c); c.input_features_dim = 0
    model, crit = build_offset_regression(c); model.init_weights(); model = model.to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.001)
    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(pts, mask, feats).transpose(1, 2), offs, mask)
        loss.backward(); torch.nn.utils.clip_grad_norm_(model.parameters(), 10); opt.step()
    t = timeit(step, n=5, warm=2)
    print("%s train step: %.2f ms -> %.2f Mpts/s" % (name, t, B * N / t / 1e3))
    def fwd():
        with torch.no_grad(): model(pts, mask, feats)
    print("%s fwd only: %.2f ms" % (name, timeit(fwd, n=5, warm=2)))
    if os.environ.get("PROFILE"):
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(); torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
