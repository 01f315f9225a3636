import torch, sys, os
sys.path.insert(0, "/root/repo")
from deep3dpointclouddenoising_b200 import ops, synthetic
dev = torch.device("cuda:0")
B, N, C, ns = 16, 8192, 72, 52
pts, mask, _, _ = [torch.from_numpy(a).to(dev) for a in synthetic.make_batch(1, B, N)]
idx = ops.ball_query(pts, pts, mask, mask, 0.025, ns)[0]
g = torch.randn(B, C, N, ns, device=dev)
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
x = ops.group_points_grad(g, idx, N, deterministic=True)
y = ops.group_points_grad(g, idx, N)
print("max rel diff", float((x - y).abs().max() / x.abs().max()))
print("deterministic %.3f ms  atomic planes %.3f ms" % (t(lambda: ops.group_points_grad(g, idx, N, deterministic=True)), t(lambda: ops.group_points_grad(g, idx, N))))
