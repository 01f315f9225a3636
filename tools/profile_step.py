"""Per-kernel time of one eager training step (torch.profiler / CUPTI), top-N table.  Diagnostic only."""
import argparse
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deep3dpointclouddenoising_b200 import synthetic  # noqa: E402
from deep3dpointclouddenoising_b200.utils.config import runtime  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--operator", default="pospool")
    ap.add_argument("--channel-last", type=int, default=1)
    ap.add_argument("--tf32", type=int, default=0)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--flat", type=int, default=0, help="flat parameter/gradient buffers + in-place parameter gradients")
    ap.add_argument("--list", default=None, help="substring: print every launch of matching kernels in order")
    args = ap.parse_args()
    runtime.channel_last = bool(args.channel_last)
    runtime.pseudo_grid_precision = args.precision
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    dev = torch.device("cuda:0")
    model, criterion, cfg = bench.build_model(args.operator, 8192)
    model = model.to(dev)
    flat = None
    if args.flat:
        from deep3dpointclouddenoising_b200 import distributed
        flat = distributed.FlatParameters(model)
        runtime.grads_in_place = True
    opt = torch.optim.Adam([flat.param] if flat else model.parameters(), lr=1e-3)
    batch = [torch.from_numpy(a).to(dev) for a in synthetic.make_batch(0, 16, 8192, ragged=True)]

    def step():
        if flat:
            flat.zero()
        else:
            opt.zero_grad(set_to_none=True)
        loss = criterion(model(batch[0], batch[1], batch[2]).transpose(1, 2), batch[3], batch[1])
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    if args.list:
        evs = [e for e in prof.events() if args.list in e.name]
        evs.sort(key=lambda e: e.time_range.start)
        print(" ".join(f"{e.name.split('::')[-1].split('(')[0][:18]}:{(e.device_time if hasattr(e, 'device_time') else e.cuda_time):.1f}"
                       for e in evs))
    rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count)
            for e in prof.key_averages()]
    rows.sort(key=lambda r: -r[1])
    total = sum(r[1] for r in rows)
    print(f"total device time {total / 1e3:.3f} ms over {sum(r[2] for r in rows)} launches")
    for name, t, n in rows[:args.top]:
        print(f"{t / 1e3:8.3f} ms  x{n:<4d} {name[:110]}")


if __name__ == "__main__":
    main()
