import os, sys
import torch
sys.path.insert(0, '.')
from deep3dpointclouddenoising_b200 import ops, synthetic
d = torch.device('cuda:0')
p, m, f, o = [torch.from_numpy(a).to(d) for a in synthetic.make_batch(1, 16, 8192)]
x = torch.randn(8192, 8192, device=d)
for _ in range(50): y = x @ x   # warm the clocks
torch.cuda.synchronize()
def t(n=20):
    for _ in range(5): ops.ball_query(p, p, m, m, 0.025, 52)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): ops.ball_query(p, p, m, m, 0.025, 52)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for rep in range(2):
    for warps in (4, 8, 16):
        for tile in (512, 1024, 2048, 4096, 8192):
            os.environ["D3D_BQ_WARPS"], os.environ["D3D_BQ_TILE"] = str(warps), str(tile)
            print(rep, warps, tile, round(t(), 3), flush=True)
