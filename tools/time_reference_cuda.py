"""Times the reference's own CUDA kernels (rebuilt for sm_100a) next to ours on the same B200: the "kernel to beat"."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import ops, synthetic
from oracle import cuda_ref

dev = torch.device("cuda:0")
def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

R = cuda_ref.RefCuda()
B, N = 16, 8192
pts, mask, feats, offs = [torch.from_numpy(a).to(dev) for a in synthetic.make_batch(1, B, N)]
out = {}
out["ball_query_8192x8192_ns52"] = {"ours_ms": timeit(lambda: ops.ball_query(pts, pts, mask, mask, 0.025, 52), 10, 3),
                                    "reference_ms": timeit(lambda: R.ball_query(pts, pts, mask, mask, 0.025, 52), 2, 1)}
out["grid_subsample_8192_to_2048"] = {"ours_ms": timeit(lambda: ops.grid_subsample(pts, mask, 2048, 0.003125), 10, 3),
                                      "reference_ms": timeit(lambda: R.grid_subsampling(pts, mask, 2048, 0.003125), 2, 1)}
sub, subm = ops.grid_subsample(pts, mask, 2048, 0.003125)
out["nearest_8192x2048"] = {"ours_ms": timeit(lambda: ops.nearest_query(pts, sub, mask, subm), 10, 3),
                            "reference_ms": timeit(lambda: R.nearest_query(pts, sub, mask, subm), 2, 1)}
idx, _ = ops.ball_query(pts, pts, mask, mask, 0.025, 52)
f = torch.randn(B, 72, N, device=dev)
out["group_points_C72"] = {"ours_ms": timeit(lambda: ops.group_points(f, idx), 5, 2),
                           "reference_ms": timeit(lambda: R.group_points(f, idx), 2, 1)}
g = torch.randn(B, 72, N, 52, device=dev)
out["group_points_grad_C72"] = {"ours_ms": timeit(lambda: ops.group_points_grad(g, idx, N), 5, 2),
                                "reference_ms": timeit(lambda: R.group_points_grad(g, idx, N), 2, 1)}
for k, v in out.items():
    v["speedup"] = v["reference_ms"] / v["ours_ms"]
    print(k, {a: round(b, 3) for a, b in v.items()})
json.dump(out, open("gpurun_out/reference_cuda_timing.json", "w"), indent=1)
