"""Chamfer distance of two 1M-point clouds (BASELINE config 5 scale): GPU grid search vs CPU KD-tree."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import ops, synthetic
from scipy.spatial import cKDTree
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
clean = synthetic.make_cloud(0, n, sigma=0.0)
noisy = (clean + np.random.default_rng(0).standard_normal((n, 3)).astype(np.float32) * 0.005).astype(np.float32)
x, y = torch.from_numpy(noisy).to(dev), torch.from_numpy(clean).to(dev)
for _ in range(2): cd = ops.chamfer_l2(x, y)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): cd = ops.chamfer_l2(x, y)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
t = time.time()
dx = cKDTree(clean).query(noisy, k=1, workers=-1)[0] ** 2
dy = cKDTree(noisy).query(clean, k=1, workers=-1)[0] ** 2
cpu_s = time.time() - t
ref = dx.mean() + dy.mean()
print(f"n={n}: GPU chamfer {ms:.2f} ms ({2 * n / ms / 1e3:.1f} Mqueries/s), value {cd[0].item():.6e}; "
      f"CPU cKDTree ({os.cpu_count()} threads) {cpu_s:.2f} s, value {ref:.6e}, rel diff {abs(cd[0].item() - ref) / ref:.2e}")
