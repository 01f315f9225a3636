"""Timeline of ONE replayed training step (the CUDA graph bench.py times), from CUPTI through torch.profiler:
per-kernel device time, per-stream busy time, idle gaps on the main stream and what the main stream was waiting for.
Diagnostic only (numbers taken under a profiler are never bench values).
usage: python tools/step_timeline.py [--operator pospool] [--top 40] [--gaps 15]"""
import argparse
import json
import os
import sys
import tempfile
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deep3dpointclouddenoising_b200 import distributed, synthetic  # noqa: E402
from deep3dpointclouddenoising_b200.utils import config as cfgmod  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--operator", default="pospool")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--gaps", type=int, default=15)
    ap.add_argument("--list", default=None, help="comma-separated substrings: print every launch of matching kernels (start, us, grid)")
    ap.add_argument("--pipeline", action="store_true", help="the one-step pipelined form (next batch's pyramid during backward)")
    args = ap.parse_args()
    for kv in filter(None, os.environ.get("D3D_RUNTIME", "").split(",")):
        k, v = kv.split("=")
        cfgmod.runtime[k] = {"0": False, "1": True}.get(v, int(v) if v.isdigit() else v)
    cfgmod.runtime.pseudo_grid_precision = args.precision
    dev = torch.device("cuda:0")
    model, criterion, cfg = bench.build_model(args.operator, 8192)
    model = model.to(dev)
    bucket = distributed.FlatParameters(model)
    cfgmod.runtime.grads_in_place = True
    opt = torch.optim.Adam([bucket.param], lr=cfg.base_learning_rate, weight_decay=cfg.weight_decay, capturable=True, fused=True)
    batch = [torch.from_numpy(a).to(dev) for a in synthetic.make_batch(0, 16, 8192)]

    nxt = [t.clone() for t in batch]

    def step(piped=False):
        from deep3dpointclouddenoising_b200 import neighbors
        bucket.zero()
        loss = criterion(model(batch[0], batch[1], batch[2]).transpose(1, 2), batch[3], batch[1])
        if piped:
            model.prefetch_neighbors(nxt[0], nxt[1])
        loss.backward()
        bucket.reduce()
        torch.nn.utils.clip_grad_norm_([bucket.param], 10)
        opt.step()
        if piped:
            neighbors.fold_pending_into_current()
            for d, src in zip(batch, nxt):
                d.copy_(src)
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    if args.pipeline:
        from deep3dpointclouddenoising_b200 import neighbors
        model.prefetch_neighbors(batch[0], batch[1])
        neighbors.settle()
    with torch.cuda.graph(graph, stream=torch.cuda.Stream(priority=-1) if args.pipeline else None):
        step(args.pipeline)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        graph.replay()
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    print(f"replayed step: {(t1 - t0) / 1e3:.3f} ms wall on the device, {len(ev)} kernels/copies, "
          f"summed kernel time {sum(e['dur'] for e in ev) / 1e3:.3f} ms")
    streams = defaultdict(list)
    for e in ev:
        streams[e["args"].get("stream", e.get("tid"))].append(e)
    main_stream = max(streams, key=lambda s: sum(e["dur"] for e in streams[s]))
    for s, lst in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
        busy = sum(e["dur"] for e in lst)
        print(f"  stream {s}: {len(lst)} launches, busy {busy / 1e3:.3f} ms, span {(lst[-1]['ts'] + lst[-1]['dur'] - lst[0]['ts']) / 1e3:.3f} ms"
              + ("  <- main" if s == main_stream else ""))
    ml = streams[main_stream]
    gaps = []
    for a, b in zip(ml, ml[1:]):
        g = b["ts"] - (a["ts"] + a["dur"])
        if g > 0:
            gaps.append((g, a["name"][:60], b["name"][:60], b["ts"] - t0))
    print(f"main stream idle between kernels: {sum(g[0] for g in gaps) / 1e3:.3f} ms in {len(gaps)} gaps "
          f"(median {sorted(g[0] for g in gaps)[len(gaps) // 2]:.1f} us)")
    for g, a, b, at in sorted(gaps, reverse=True)[:args.gaps]:
        print(f"    {g:7.1f} us at +{at / 1e3:6.3f} ms  after {a}  before {b}")
    if args.list:
        for pat in args.list.split(","):
            sel = [e for e in ev if pat in e["name"]]
            print(f"launches of *{pat}*: " + " ".join(f"{(e['ts'] - t0) / 1e3:.2f}:{e['dur']:.1f}us:{e['args'].get('grid', '')}" for e in sel))
    # exposure: time during which a kernel runs ALONE (nothing else on the device) — what shortening it would save — and
    # the time the device sits idle between kernels
    points = sorted({e["ts"] for e in ev} | {e["ts"] + e["dur"] for e in ev})
    alone = defaultdict(float)
    idle = 0.0
    starts = sorted(ev, key=lambda e: e["ts"])
    active, si = [], 0
    for a, b in zip(points, points[1:]):
        while si < len(starts) and starts[si]["ts"] <= a:
            active.append(starts[si])
            si += 1
        active = [e for e in active if e["ts"] + e["dur"] > a]
        if not active:
            idle += b - a
        elif len(active) == 1:
            alone[active[0]["name"]] += b - a
    print(f"device idle inside the step: {idle / 1e3:.3f} ms; time with exactly one kernel running: {sum(alone.values()) / 1e3:.3f} ms")
    agg = defaultdict(lambda: [0.0, 0, 0.0])
    for e in ev:
        k = e["name"].replace("(anonymous namespace)::", "").replace("void ", "")[:90]
        agg[k][0] += e["dur"]
        agg[k][1] += 1
    for name, d in alone.items():
        agg[name.replace("(anonymous namespace)::", "").replace("void ", "")[:90]][2] += d
    print("per kernel (this replay): total, launches, of which running alone")
    for k, (d, n, al) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:args.top]:
        print(f"  {d / 1e3:7.3f} ms x{n:<4d} alone {al / 1e3:6.3f} ms  {k}")


if __name__ == "__main__":
    main()
