"""Experiment: capture the whole training step (forward, loss, backward, clip, Adam) in one CUDA graph."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deep3dpointclouddenoising_b200 import synthetic
dev = torch.device("cuda:0")
B, N = 16, 8192
model, criterion, cfg = bench.build_model("pospool", N)
model = model.to(dev)
opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.001, capturable=True)
batches = [[torch.from_numpy(a).to(dev) for a in synthetic.make_batch(10 + s, B, N)] for s in range(4)]
static = [t.clone() for t in batches[0]]

def step():
    opt.zero_grad(set_to_none=False)
    loss = criterion(model(static[0], static[1], static[2]).transpose(1, 2), static[3], static[1])
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10)
    opt.step()
    return loss

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("eager step: %.2f ms" % timeit(step))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    static_loss = step()
torch.cuda.synchronize()
def replay():
    g.replay()
print("graph replay: %.2f ms" % timeit(replay))
l0 = None
for k in range(8):
    for dst, src in zip(static, batches[k % 4]): dst.copy_(src)
    g.replay()
    print("loss", static_loss.item())
