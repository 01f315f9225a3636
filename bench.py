#!/usr/bin/env python
"""bench.py — training points/s of the point U-Net offset-regression step on synthetic PointCleanNet-shaped patches.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                      (the reference algorithm on the host cores, rank 0 only)

Workload (BASELINE.json configs[1]): cfgs/l1.yaml geometry as train_dist.py imposes it, PosPool aggregation
(`--operator pseudo_grid` for configs[2]), batch 16 x 8192-point patches per GPU, fp32, Adam, grad-clip 10.
A step = forward + MaskedL1 loss + backward + clip + optimizer step (train_dist.py:440-451).

One JSON line on rank 0:
  value        points/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e          same metric through the public API with HOST batches: pinned H2D copy of every step's inputs and a
               D2H read of the loss inside the timed region
  roofline     dominant kernel of the step: algorithmic bytes / CUDA-event duration measured live vs MEASURED_PEAKS.json
  cpu_baseline the oracle port of the same step (oracle/cpu_model.py) on the host cores, bounded sample
  kernels      per-op device time inside one instrumented step (explains value; not part of the contract)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_points_per_sec_fwd_bwd"
UNIT = "points/s"
BATCH, NUM_POINTS = 16, 8192
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--operator", default="pospool", choices=["pospool", "pseudo_grid"])
    p.add_argument("--pseudo-grid-precision", default="fp32", choices=["fp32", "bf16"])
    p.add_argument("--batch", type=int, default=BATCH)
    p.add_argument("--num-points", type=int, default=NUM_POINTS)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extras", action="store_true",
                   help="skip the other BASELINE configs (PseudoGrid step, 100k pyramid, 1M-point inference) and the "
                        "reference-GPU-path timing that the default 1-GPU run appends under 'extra'")
    p.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                   help="replay the whole training step as one CUDA graph (auto: on)")
    return p.parse_args()


def build_model(operator, num_points):
    from deep3dpointclouddenoising_b200.models import build_offset_regression
    from deep3dpointclouddenoising_b200.utils import config as cfgmod
    cfgmod.reset_config()
    name = "l1.yaml" if operator == "pseudo_grid" else "l1_pospool.yaml"
    cfgmod.update_config(os.path.join(ROOT, "deep3dpointclouddenoising_b200", "cfgs", name))
    c = cfgmod.config
    c.num_points = num_points
    cfgmod.apply_train_geometry(c)
    c.input_features_dim = 0
    torch.manual_seed(0)
    np.random.seed(0)
    model, criterion = build_offset_regression(c)
    model.init_weights()
    return model, criterion, c


def workload_name(operator, batch, num_points):
    return (f"U-Net offset regression fwd+bwd, cfgs/l1.yaml geometry (train_dist.py:125-137), "
            f"{'PosPool xyz/avg' if operator == 'pospool' else 'PseudoGrid 15 kernel points'}, "
            f"batch {batch} x {num_points}-point synthetic noisy patches per GPU")


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# algorithmic bytes of one call of each op (compulsory HBM traffic, fp32/int32)
# ------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(op, s):
    B, M, N, C, ns = s.get("B", 0), s.get("M", 0), s.get("N", 0), s.get("C", 0), s.get("ns", 0)
    if op == "ball_query":
        return 16 * B * (M + N) + 8 * B * M * ns
    if op == "nearest_query":
        return 16 * B * (M + N) + 8 * B * M
    if op == "grid_subsample":
        return 16 * B * N + 16 * B * M
    if op == "build_inverse_map":
        return 4 * B * M * ns * 2 + 4 * B * N
    if op in ("cm_to_cl", "cl_to_cm"):
        return 8 * B * C * N
    if op in ("pospool_fwd", "pseudogrid_fwd", "gather_max_fwd"):
        return 4 * B * C * N + 4 * B * M * ns + 4 * B * M + 12 * B * (M + N) + 4 * B * C * M
    if op in ("pospool_bwd", "pseudogrid_bwd", "gather_max_bwd"):
        return 4 * B * C * M + 4 * B * M * ns + 4 * B * N + 12 * B * (M + N) + 4 * B * C * N
    if op in ("nearest_gather_fwd", "nearest_gather_bwd"):
        return 4 * B * C * (M + N) + 4 * B * M
    R = s.get("R", 0)
    if op == "bn_act_cl_fwd":   # two passes over x (statistics, apply) + y out (+ residual in): minimum of a standalone BN
        return 4 * R * C * (3 + s.get("res", 0))
    if op == "bn_from_stats":   # statistics came from the GEMM epilogue: one pass, x in, y out (+ residual in)
        return 4 * R * C * (2 + s.get("res", 0))
    if op == "bn_act_cl_bwd":   # dy and x twice (reduce, apply) + dx out (+ y in for the ReLU mask, + dres out)
        return 4 * R * C * (5 + s.get("y", 0) + s.get("res", 0))
    return 0


def algorithmic_flops(op, s):
    """1x1 convolutions as GEMMs: forward 2 R Cin Cout; backward = data gradient + weight gradient."""
    R, ci, co = s.get("R", 0), s.get("Cin", 0), s.get("Cout", 0)
    if op == "conv1x1_fwd":
        return 2 * R * ci * co
    if op == "conv1x1_bwd":
        return 4 * R * ci * co
    return 0


class OpTimer:
    """Wraps the ops module's functions with CUDA-event pairs on the current stream (instrumented pass only)."""

    def __init__(self, ops):
        self.ops, self.records, self.saved = ops, [], {}

    def _shape(self, name, args):
        t = [a for a in args if isinstance(a, torch.Tensor)]
        try:
            if name == "ball_query":
                return dict(B=t[0].shape[0], M=t[0].shape[1], N=t[1].shape[1], ns=int(args[5]))
            if name == "nearest_query":
                return dict(B=t[0].shape[0], M=t[0].shape[1], N=t[1].shape[1])
            if name == "grid_subsample":
                return dict(B=t[0].shape[0], N=t[0].shape[1], M=int(args[2]))
            if name == "build_inverse_map":
                i = t[0]
                return dict(B=i.shape[0], M=i.shape[1], ns=i.shape[2] if i.dim() == 3 else 1, N=int(args[1]))
            if name == "cm_to_cl":
                return dict(B=t[0].shape[0], C=t[0].shape[1], N=t[0].shape[2])
            if name == "cl_to_cm":
                return dict(B=t[0].shape[0], C=t[0].shape[2], N=t[0].shape[1])
            if name in ("pospool_fwd", "pseudogrid_fwd"):
                return dict(B=t[0].shape[0], N=t[0].shape[1], C=t[0].shape[2], M=t[3].shape[1], ns=t[3].shape[2])
            if name == "gather_max_fwd":
                return dict(B=t[0].shape[0], N=t[0].shape[1], C=t[0].shape[2], M=t[1].shape[1], ns=t[1].shape[2])
            if name == "pospool_bwd":
                return dict(B=t[0].shape[0], M=t[0].shape[1], C=t[0].shape[2], N=int(args[7]), ns=int(args[8]))
            if name == "pseudogrid_bwd":
                return dict(B=t[0].shape[0], M=t[0].shape[1], C=t[0].shape[2], N=t[1].shape[1], ns=t[4].shape[2])
            if name == "gather_max_bwd":
                return dict(B=t[0].shape[0], M=t[0].shape[1], C=t[0].shape[2], N=int(args[4]), ns=0)
            if name == "nearest_gather_fwd":
                return dict(B=t[0].shape[0], N=t[0].shape[1], C=t[0].shape[2], M=t[1].shape[1])
            if name == "nearest_gather_bwd":
                return dict(B=t[0].shape[0], M=t[0].shape[1], C=t[0].shape[2], N=int(args[3]))
            if name == "bn_act_cl_fwd":
                x = args[0]
                return dict(R=x.numel() // x.shape[-1], C=x.shape[-1], res=0 if args[1] is None else 1)
            if name == "bn_from_stats":
                x = args[0]
                return dict(R=x.numel() // x.shape[-1], C=x.shape[-1], res=0 if args[1] is None else 1)
            if name == "bn_act_cl_bwd":
                x = args[0]
                return dict(R=x.numel() // x.shape[-1], C=x.shape[-1], y=0 if args[2] is None else 1, res=1 if args[9] else 0)
        except Exception:
            pass
        return {}

    def __enter__(self):
        names = ["ball_query", "nearest_query", "grid_subsample", "build_inverse_map", "cm_to_cl", "cl_to_cm",
                 "pospool_fwd", "pospool_bwd", "pseudogrid_fwd", "pseudogrid_bwd", "gather_max_fwd", "gather_max_bwd",
                 "nearest_gather_fwd", "nearest_gather_bwd", "group_points", "group_points_grad", "spatial_order",
                 "bn_act_cl_fwd", "bn_act_cl_bwd", "bn_from_stats"]
        for n in names:
            fn = getattr(self.ops, n)
            self.saved[n] = fn

            def wrapped(*a, _fn=fn, _n=n, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                out = _fn(*a, **k)
                e.record()
                self.records.append((_n, self._shape(_n, a), s, e))
                return out

            setattr(self.ops, n, wrapped)
        # the 1x1 convolutions (row-major GEMMs, models/blocks.py): forward / backward of the two autograd Functions
        from deep3dpointclouddenoising_b200.models import blocks
        self.conv_saved = []
        for cls in (blocks.PointwiseConvRows, blocks.PointwiseConvCatRows):
            for meth, tag in (("forward", "conv1x1_fwd"), ("backward", "conv1x1_bwd")):
                fn = getattr(cls, meth)
                self.conv_saved.append((cls, meth, fn))

                def timed(ctx, *a, _fn=fn, _tag=tag, _cat=cls is blocks.PointwiseConvCatRows, _fwd=meth == "forward"):
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    out = _fn(ctx, *a)
                    e.record()
                    if _fwd:
                        w = a[0] if _cat else a[1]
                        rows = a[3:] if _cat else a[:1]  # cat: (weight, bias, want_stats, *rows)
                    else:
                        saved = ctx.saved_tensors
                        w = saved[0] if _cat else saved[1]
                        rows = saved[1:] if _cat else saved[:1]
                    shape = dict(R=rows[0].numel() // rows[0].shape[-1], Cin=sum(r.shape[-1] for r in rows), Cout=w.shape[0])
                    self.records.append((_tag, shape, s, e))
                    return out

                setattr(cls, meth, staticmethod(timed))
        return self

    def __exit__(self, *exc):
        for n, fn in self.saved.items():
            setattr(self.ops, n, fn)
        for cls, meth, fn in self.conv_saved:
            setattr(cls, meth, staticmethod(fn))

    def summary(self, n_steps):
        torch.cuda.synchronize()
        per = {}
        for name, shape, s, e in self.records:
            key = name + "".join(f" {k}{v}" for k, v in sorted(shape.items()))
            d = per.setdefault(key, {"op": name, "shape": shape, "times": [], "calls": 0})
            d["times"].append(s.elapsed_time(e))
            d["calls"] += 1
        for d in per.values():
            # median over the calls of a shape: one call that happened to queue behind another stream's work (the neighbour
            # ops run on a side stream in this eager pass) must not be charged to the op
            d["ms_per_call"] = float(np.median(d.pop("times")))
            d["ms"] = d["ms_per_call"] * d["calls"]
            d["ms_per_step"] = d["ms"] / n_steps
            d["bytes_per_call"] = algorithmic_bytes(d["op"], d["shape"])
            d["flops_per_call"] = algorithmic_flops(d["op"], d["shape"])
        return per


# ------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the same step)
# ------------------------------------------------------------------------------------------------------------
def cpu_step_fn(operator, num_points):
    from oracle import cpu_index_ops
    from oracle.cpu_model import CpuUNet
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    torch.set_num_threads(os.cpu_count() or 1)
    model, criterion, cfg = build_model(operator, num_points)
    opt = torch.optim.Adam(model.parameters(), lr=cfg.base_learning_rate, weight_decay=cfg.weight_decay)
    # index ops: the reference's own kernels compiled for the host when the prebuilt library travelled, else the port
    use_ref = cpu_index_ops.reference_available()
    ix = cpu_index_ops.reference() if use_ref else cpu_index_ops.restated()
    net = CpuUNet(model, ix)

    def step(batch):
        pts, mask, feats, offs = [torch.from_numpy(a) for a in batch]
        opt.zero_grad(set_to_none=True)
        loss = criterion(net(pts, mask, feats).transpose(1, 2), offs, mask)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10)
        opt.step()
        return float(loss)

    kind = "port"  # aggregation math is the oracle restatement even when the index kernels are the reference's
    return step, kind, ("reference kernels (host build) + " if use_ref else "") + "oracle port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from deep3dpointclouddenoising_b200 import synthetic
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    # the GPU arm's batch when the host can afford it (one 16 x 8192 step costs 5-10 s on 16 cores), else one patch per
    # step (the metric is per point and patches are independent; BatchNorm statistics are then per patch);
    # the whole run stays within a few minutes
    num_points = args.num_points
    budget_s = 240.0
    step, kind, how = cpu_step_fn(args.operator, num_points)
    t0 = time.time()
    step(synthetic.make_batch(999, 1, num_points))
    probe1 = time.time() - t0
    patches = args.batch if probe1 * args.batch * 3 <= budget_s else 1
    probe = probe1 * patches
    steps, warm = args.steps, args.warmup
    if probe * total > budget_s:  # shrink what a "step" executes, never the patch geometry
        warm = min(warm, 1)
        steps = max(1, min(steps, int(budget_s / probe) - warm))
    for w in range(warm):
        step(synthetic.make_batch(1000 + w, patches, num_points))
    t0 = time.time()
    for s in range(steps):
        step(synthetic.make_batch(2000 + s, patches, num_points))
    dt = time.time() - t0
    value = patches * num_points * steps / dt
    sample = (f"{steps} timed steps (+{warm} warm-up) of {patches} x {num_points}-point patch fwd+bwd+Adam on {cores} host threads "
              f"({how}); requested steps={args.steps} warmup={args.warmup} bounded to ~{int(budget_s)} s")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.operator, args.batch, args.num_points)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# the other BASELINE configs and the reference's GPU path, appended to the same JSON line under "extra"
# ------------------------------------------------------------------------------------------------------------
def reference_gpu_step(args, dev, batch):
    """The reference's own GPU path on this B200 (oracle/gpu_reference.py): its CUDA kernels rebuilt for sm_100a, its
    eager aggregation, stock cuDNN / ATen modules — one full training step at the bench batch."""
    from oracle import cuda_ref, gpu_reference
    from deep3dpointclouddenoising_b200.utils import config as cfgmod
    if not cuda_ref.available():
        return {"unavailable": "oracle/_ref/libref_cuda.so did not travel (built where /root/reference exists)"}
    rt = cfgmod.runtime
    saved = dict(rt)
    rt.fused_batchnorm, rt.channel_last, rt.prefetch_neighbors, rt.grads_in_place = False, False, False, False
    try:
        model, criterion, cfg = build_model(args.operator, args.num_points)
        model = model.to(dev)
        step = gpu_reference.make_step(model, criterion, cfg.base_learning_rate, cfg.weight_decay)
        torch.cuda.reset_peak_memory_stats(dev)
        step(*batch)  # warm-up (cuDNN autotune, allocator)
        torch.cuda.synchronize()
        n = 2
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            step(*batch)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / n
        B, N = batch[0].shape[0], batch[0].shape[1]
        out = {"value": B * N / (ms / 1e3), "unit": UNIT, "ms_per_step": round(ms, 2), "steps": n, "batch": B,
               "num_points": N, "peak_memory_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2),
               "what": "reference CUDA kernels (unmodified, rebuilt for sm_100a: one ball query per call, materialised "
                       "gathers, atomicAdd scatter) + the reference's eager PyTorch aggregation and stock Conv1d / "
                       "BatchNorm1d modules (TF32 convolutions), Adam; Python glue = the oracle port of the reference's "
                       "call sequence (oracle/gpu_reference.py)"}
        del model, step
        return out
    finally:
        for k, v in saved.items():
            rt[k] = v
        torch.cuda.empty_cache()


def run_extras(args, dev, batch):
    extra = {}
    try:
        extra["reference_gpu"] = reference_gpu_step(args, dev, batch)
    except Exception as e:  # never lose the headline line to an extra
        extra["reference_gpu"] = {"error": f"{type(e).__name__}: {e}"}
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    torch.cuda.empty_cache()  # the training bench leaves several GB cached in its own block sizes
    try:
        import pyramid_100k
        extra["config1_pyramid_100k"] = pyramid_100k.run(dev, n_rep=5, cpu=True, single_thread=False)
    except Exception as e:
        extra["config1_pyramid_100k"] = {"error": f"{type(e).__name__}: {e}"}
    try:
        import denoise_1m
        extra["config5_inference_1m"] = denoise_1m.run(dev, 1_000_000, 8192, cpu=True)
    except Exception as e:
        extra["config5_inference_1m"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    if args.operator == "pospool":  # BASELINE config 3: the PseudoGrid step, both arithmetic modes, in a process of its own
        for prec in ("bf16", "fp32"):
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--operator", "pseudo_grid",
                                    "--pseudo-grid-precision", prec, "--steps", "30", "--warmup", "5", "--no-extras",
                                    "--no-cpu-baseline", "--batch", str(args.batch), "--num-points", str(args.num_points)],
                                   capture_output=True, text=True, timeout=400)
                line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
                extra[f"config3_pseudogrid_{prec}"] = {
                    "value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"], "steps": line["steps"],
                    "e2e": line["e2e"]["value"], "workload": line["config"]["workload"],
                    "top_kernels": dict(list(line["families"].items())[:4])}
            except Exception as e:
                extra[f"config3_pseudogrid_{prec}"] = {"error": f"{type(e).__name__}: {e}"}
    return extra


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    from deep3dpointclouddenoising_b200 import _lib, distributed, ops, synthetic
    from deep3dpointclouddenoising_b200.utils import config as cfgmod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.graph != "off":
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")  # required for capturing NCCL collectives
    rank, world, local_rank = distributed.init("nccl")
    lib = _lib.load()
    cfgmod.runtime.pseudo_grid_precision = args.pseudo_grid_precision
    for kv in filter(None, os.environ.get("D3D_RUNTIME", "").split(",")):  # experiments: D3D_RUNTIME=own_gemm=0,staged_tiles=0
        k, v = kv.split("=")
        cfgmod.runtime[k] = {"0": False, "1": True}.get(v, int(v) if v.isdigit() else v)

    model, criterion, cfg = build_model(args.operator, args.num_points)
    model = model.to(dev)
    # gradient averaging when world > 1 (train_dist.py:375 uses DDP): `--graph off` wraps the model in DDP exactly like
    # the reference; the graph mode uses one flat gradient bucket + a single NCCL all-reduce that is captured with
    # the rest of the step (DDP's reducer hooks do not replay)
    # In graph mode parameters and gradients are additionally views of two flat buffers (distributed.FlatParameters):
    # Adam and the gradient clipping then run on one tensor — same arithmetic, a handful of launches.
    bucket = None
    use_graph = args.graph != "off"
    if use_graph:
        net = model
        bucket = distributed.FlatParameters(model)
        if world > 1 and os.environ.get("D3D_ALLREDUCE_OVERLAP", "1") != "0":
            bucket.overlap_with_backward(model)  # slices of the bucket go out while backward still runs
        opt_params = [bucket.param]
        # every gradient buffer exists for the whole run and nothing hooks the gradients: the backward kernels add
        # into the buffers directly instead of returning tensors that autograd accumulates with ~120 tiny add kernels
        cfgmod.runtime.grads_in_place = True
    else:
        net = distributed.wrap(model, local_rank)
        opt_params = list(model.parameters())
    # fused=True: torch's single-kernel Adam (same update rule) — with the flat parameter it is ONE launch per step
    opt = torch.optim.Adam(opt_params, lr=cfg.base_learning_rate, weight_decay=cfg.weight_decay,
                           capturable=use_graph, fused=True if use_graph else None)
    B, N = args.batch, args.num_points

    def host_batch(step):  # per-rank shard of the synthetic patch stream
        arrs = synthetic.make_batch(distributed.shard_seed(rank, step), B, N)
        return [torch.from_numpy(a).pin_memory() for a in arrs]

    def train_step(pts, mask, feats, offs):
        if bucket is not None:
            bucket.zero()
        else:
            opt.zero_grad(set_to_none=True)
        loss = criterion(net(pts, mask, feats).transpose(1, 2), offs, mask)
        loss.backward()
        if bucket is not None:
            bucket.reduce()
        torch.nn.utils.clip_grad_norm_(opt_params, 10)
        opt.step()
        return loss

    # one-step software pipeline (graph mode): the neighbourhood pyramid of batch k+1 is built on the side stream while
    # the backward pass of batch k runs (it depends on coordinates only), and handed over at fixed addresses at the end
    # of the step; every replay trains on the batch the previous replay received.  D3D_PIPELINE=0: build inside forward.
    pipelined = use_graph and os.environ.get("D3D_PIPELINE", "1") != "0"
    handover = torch.cuda.Stream() if pipelined else None

    def train_step_pipelined(cur, nxt):
        from deep3dpointclouddenoising_b200 import neighbors
        bucket.zero()
        loss = criterion(net(cur[0], cur[1], cur[2]).transpose(1, 2), cur[3], cur[1])
        model.prefetch_neighbors(nxt[0], nxt[1])
        loss.backward()
        bucket.reduce()
        # hand-over on a second stream, concurrently with the clipping and the optimiser (backward is complete: nothing
        # reads this batch's pyramid or inputs any more)
        main = torch.cuda.current_stream()
        handover.wait_stream(main)
        with torch.cuda.stream(handover):
            neighbors.fold_pending_into_current()  # next batch's pyramid -> the addresses the next replay's forward reads
            for d, src in zip(cur, nxt):
                d.copy_(src)
        torch.nn.utils.clip_grad_norm_(opt_params, 10)
        opt.step()
        main.wait_stream(handover)
        return loss

    def barrier():
        distributed.barrier(dev)

    n_host = 4
    host = [host_batch(s) for s in range(n_host)]
    resident = [[t.to(dev) for t in hb] for hb in host]
    static = [t.clone() for t in resident[0]]  # the step always reads these buffers (CUDA-graph friendly)
    static_next = [t.clone() for t in resident[0]] if pipelined else None  # where the next batch lands

    # ---- warm-up (eager, on a side stream as graph capture requires), launch count of one step, capture ----
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for w in range(max(args.warmup, 3, 11 if (use_graph and world > 1) else 0)):
            for d, src in zip(static, resident[w % n_host]):
                d.copy_(src)
            if w == 2:
                launches0 = lib.d3d_kernel_launches()
            train_step(*static)
            if w == 2:
                launches_per_step = lib.d3d_kernel_launches() - launches0
    torch.cuda.current_stream().wait_stream(side)
    barrier()
    graph, static_loss = None, None
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        if pipelined:
            from deep3dpointclouddenoising_b200 import neighbors
            model.prefetch_neighbors(static[0], static[1])  # the pyramid the first replay's forward finds
            neighbors.settle()
            # the step's own kernels are captured at high priority: the prefetch (side stream, default priority) fills the
            # SM slots they leave free instead of queueing ahead of them
            hi = torch.cuda.Stream(priority=-1) if os.environ.get("D3D_PIPELINE_PRIORITY", "1") != "0" else None
            with torch.cuda.graph(graph, stream=hi):
                static_loss = train_step_pipelined(static, static_next)
        else:
            with torch.cuda.graph(graph):
                static_loss = train_step(*static)
        barrier()

    def run_step(batch, non_blocking=False):
        for d, src in zip(static_next if (pipelined and graph is not None) else static, batch):
            d.copy_(src, non_blocking=non_blocking)
        if graph is not None:
            graph.replay()
            return static_loss
        return train_step(*static)

    # ---- device-resident timing -------------------------------------------------------------------------
    for w in range(3):
        run_step(resident[w % n_host])
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for s in range(args.steps):
        run_step(resident[s % n_host])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    clock_info = clocks.stop()

    # ---- end to end: host batches in, loss out ----------------------------------------------------------
    def e2e_step(hb):
        return run_step(hb, non_blocking=True).item()  # pinned H2D copies of the inputs, D2H read of the loss

    for w in range(3):
        e2e_step(host[w % n_host])
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(args.steps):
        e2e_step(host[s % n_host])
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    ms, ms_e2e = distributed.max_over_ranks([ms, ms_e2e], dev)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # ---- instrumented pass: where the step's device time goes, and the roofline of the dominant kernel ----
    # (every rank runs these steps — they contain the gradient all-reduce — only rank 0 reports them)
    kernels, roofline, families = None, None, None
    n_inst = 3
    for s in range(2):  # eager again after graph replay: let the caching allocator settle before timing single ops
        train_step(*resident[s % n_host])
    barrier()
    with OpTimer(ops) as timer:
        for s in range(n_inst):
            # keep the GPU busy while the host enqueues the whole eager step: otherwise an op made of several launches
            # (BatchNorm = statistics + apply, ball query = grid + nearest + fill) is charged the host's launch gaps
            torch.cuda._sleep(int(8e7))
            train_step(*resident[s % n_host])
        per = timer.summary(n_inst)
    barrier()
    rooflines = None
    if rank == 0:
        step_ms = ms / args.steps
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            pk = json.load(open(peaks_path))
            peak, which = float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
            # kind::tf32 runs at half the bf16 rate: half the measured sustained bf16 figure (the GEMMs run inside a long step)
            tf32_peak, tf32_which = float(pk["bf16_tflops_sustained"]) / 2, "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32)"
        else:
            peak, which = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
            tf32_peak, tf32_which = 1400.0 / 2, "fallback sustained bf16 / 2 (tf32)"
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        traffic_table = json.load(open(tpath)) if os.path.exists(tpath) else {}

        def roof(key, v):
            if v["flops_per_call"]:
                ach = v["flops_per_call"] / 1e9 / v["ms_per_call"]  # TFLOP/s
                r = {"bound": "tensor", "achieved": round(ach, 2), "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": round(ach / tf32_peak, 4), "peak_source": tf32_which,
                     "algorithmic_flops_per_launch": v["flops_per_call"]}
            else:
                ach = v["bytes_per_call"] / 1e6 / v["ms_per_call"]  # GB/s
                r = {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                     "peak_source": which, "algorithmic_bytes_per_launch": v["bytes_per_call"]}
            r.update({"kernel": key, "traffic": traffic_table.get(key), "ms_per_launch": round(v["ms_per_call"], 4),
                      "launches_per_step": v["calls"] / n_inst, "share_of_step": round(v["ms_per_step"] / step_ms, 4)})
            return r

        ranked = sorted(per.items(), key=lambda kv: -kv[1]["ms"])
        kernels = {k: {"ms_per_step": round(v["ms_per_step"], 4), "calls_per_step": v["calls"] / n_inst,
                       "gbs": round(v["bytes_per_call"] / 1e6 / v["ms_per_call"], 1) if v["ms_per_call"] > 0 and v["bytes_per_call"] else None,
                       "tflops": round(v["flops_per_call"] / 1e9 / v["ms_per_call"], 2) if v["flops_per_call"] else None}
                   for k, v in ranked[:16]}
        ours_ms = sum(v["ms_per_step"] for v in per.values() if not v["op"].startswith("conv1x1"))
        # aggregate by op family as well: the step launches the same kernel at five resolutions
        fam = {}
        for k, v in per.items():
            f = fam.setdefault(v["op"], {"ms_per_step": 0.0, "bytes": 0.0, "flops": 0.0, "launches": 0.0})
            f["ms_per_step"] += v["ms_per_step"]
            f["bytes"] += v["bytes_per_call"] * v["calls"] / n_inst
            f["flops"] += v["flops_per_call"] * v["calls"] / n_inst
            f["launches"] += v["calls"] / n_inst
        families = {k: {"ms_per_step": round(f["ms_per_step"], 4), "launches_per_step": f["launches"],
                        "share_of_step": round(f["ms_per_step"] / step_ms, 4),
                        "gbs": round(f["bytes"] / 1e6 / f["ms_per_step"], 1) if f["bytes"] and f["ms_per_step"] > 0 else None,
                        "frac_of_hbm_peak": round(f["bytes"] / 1e6 / f["ms_per_step"] / peak, 4) if f["bytes"] and f["ms_per_step"] > 0 else None,
                        "tflops": round(f["flops"] / 1e9 / f["ms_per_step"], 2) if f["flops"] and f["ms_per_step"] > 0 else None,
                        "frac_of_tf32_peak": round(f["flops"] / 1e9 / f["ms_per_step"] / tf32_peak, 4) if f["flops"] and f["ms_per_step"] > 0 else None}
                    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms_per_step"])}
        # `roofline`: the dominant KERNEL FAMILY of the step — one op over all its launches (the step runs every op at five
        # resolutions): achieved = summed algorithmic bytes (or flops) / summed device time = average bytes per launch /
        # average launch duration.  `rooflines`: the five largest families; `per_shape`: the five largest single shapes.
        def fam_roof(name, f):
            n = max(f["launches"], 1e-9)
            v = {"flops_per_call": f["flops"] / n, "bytes_per_call": f["bytes"] / n, "ms_per_call": f["ms_per_step"] / n,
                 "calls": f["launches"] * n_inst, "ms_per_step": f["ms_per_step"]}
            r = roof(name, v)
            r["kernel"] = f"{name} (all {f['launches']:g} launches of the step, five resolutions)"
            shapes = [k for k, pv in per.items() if pv["op"] == name]
            known = [traffic_table.get(k) for k in shapes]
            r["traffic"] = None  # per-shape ncu DRAM traffic: see `per_shape` and profiles/roofline_traffic.json
            r["traffic_largest_shape"] = next((t for t in known if t), None)
            return r

        fam_ranked = [(k, f) for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms_per_step"])
                      if f["ms_per_step"] > 0 and (f["bytes"] or f["flops"])]
        rooflines = [fam_roof(k, f) for k, f in fam_ranked[:5]]
        per_shape = [roof(k, v) for k, v in ranked[:5] if v["ms_per_call"] > 0]
        roofline = dict(rooflines[0])
        roofline["per_shape"] = per_shape
        roofline["our_kernels_share_of_step"] = round(ours_ms / step_ms, 4)
        roofline["note"] = ("dominant kernel family of the step by summed device time (CUDA events around every C-ABI call, every "
                            "fused BatchNorm and every 1x1-convolution GEMM of an instrumented eager pass; neighbourhood ops run on "
                            "a side stream there, so their times include contention); algorithmic bytes = compulsory HBM traffic "
                            "(DESIGN.md §3); the ordered ball query is instruction-issue bound (67 % of issue slots, "
                            "profiles/r01_ball_query_v3_ncu.md): its HBM fraction says little; ncu evidence under profiles/")

    # ---- neighbour build alone (BASELINE.json metric, second figure): the full 5-level pyramid of one batch ----
    neighbor_build = None
    if rank == 0:
        from deep3dpointclouddenoising_b200 import ops as d3d_ops

        def pyramid(pts, mask):  # resnet.py:94-188 + multi_dimensional_head.py:62-85: 4 subsamplings, 9 distinct ball
            xyz, m, r, dl = pts, mask, float(cfg.radius), float(cfg.sampleDl)  # queries, 4 nearest-upsample queries
            d3d_ops.ball_query(xyz, xyz, m, m, r, cfg.nsamples[0], want_nvalid=True)
            levels = [(xyz, m)]
            for stage in range(4):
                dl *= 2
                sx, sm = d3d_ops.grid_subsample(xyz, m, cfg.npoints[stage], dl)
                d3d_ops.ball_query(sx, xyz, sm, m, r, cfg.nsamples[stage], want_nvalid=True)
                r *= 2
                d3d_ops.ball_query(sx, sx, sm, sm, r, cfg.nsamples[stage + 1], want_nvalid=True)
                levels.append((sx, sm))
                xyz, m = sx, sm
            for stage in range(4, 0, -1):
                d3d_ops.nearest_query(levels[stage - 1][0], levels[stage][0], levels[stage - 1][1], levels[stage][1])
            return sum(lv[0].shape[1] for lv in levels)

        for _ in range(3):
            pyramid(resident[0][0], resident[0][1])
        n_rep = 20
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for s in range(n_rep):
            pyramid(resident[s % n_host][0], resident[s % n_host][1])
        t1.record()
        torch.cuda.synchronize()
        nb_ms = t0.elapsed_time(t1) / n_rep
        neighbor_build = {"value": round(B * N / nb_ms / 1e3, 2), "unit": "Mpts/s (patch points per second, one GPU)",
                          "ms_per_batch": round(nb_ms, 4),
                          # ball queries: N + 2 per coarser level; nearest queries: every level but the coarsest
                          "query_points_per_batch": B * (N + 2 * sum(cfg.npoints) + N + sum(cfg.npoints[:3])),
                          "ops": "4 grid subsamplings + 9 ordered ball queries + 4 nearest queries, eager launches"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        step_cpu, kind, how = cpu_step_fn(args.operator, N)
        t, n_cpu = time.time(), 0
        while n_cpu < 64 and (time.time() - t < 12.0 or n_cpu < 2):  # ~10-30 s of CPU work, at the GPU arm's batch
            step_cpu(synthetic.make_batch(4242 + n_cpu, B, N))
            n_cpu += 1
        dt = time.time() - t
        cpu_baseline = {"value": n_cpu * B * N / dt, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{n_cpu} steps of {B} x {N}-point patches fwd+bwd+Adam ({how}), {dt:.1f} s on {cores} host threads"}

    extra = None
    if rank == 0 and world == 1 and not args.no_extras:
        extra = run_extras(args, dev, resident[0])

    if rank == 0:
        pts_per_step = world * B * N
        line = {"metric": METRIC, "value": pts_per_step * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args.operator, B, N), "global_batch": world * B,
                           "num_points": N, "operator": args.operator,
                           "pseudo_grid_precision": args.pseudo_grid_precision if args.operator == "pseudo_grid" else None,
                           "parallelism": f"dp{world}", "optimizer": "adam",
                           "grad_allreduce": None if world == 1 else (
                               ("flat bucket, 5 NCCL all-reduce slices overlapped with backward" if getattr(bucket, "_comm", None) is not None
                                else "flat bucket, 1 NCCL all-reduce") if bucket is not None else "DDP"),
                           "cuda_graph": graph is not None,
                           "pyramid": ("next batch's neighbourhood pyramid built during this batch's backward (one-step pipeline, "
                                       "one pyramid per step)") if (pipelined and graph is not None) else "built inside forward on a side stream",
                           "pospool_backward": {"scatter": "tensor-core tiles, partial sums added with float atomics like the reference's backward",
                                                "ordered": "tensor-core tiles + fixed-order second pass (no float atomics)",
                                                False: "segmented reduction over the inverse map (no float atomics)"}.get(
                               cfgmod.runtime.staged_tiles_backward, str(cfgmod.runtime.staged_tiles_backward))
                           if args.operator == "pospool" else None,
                           "layout": "channel-last activations end to end" if cfgmod.runtime.channel_last else "channel-major",
                           "conv_math": "tf32" if torch.backends.cudnn.allow_tf32 else "fp32", "l2": "per-step working set (activations, "
                           "several GB) exceeds the 126 MB L2; 4 rotating input batches"},
                "e2e": {"value": pts_per_step * args.steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline, "rooflines": rooflines,
                "families": families, "kernels": kernels}
        if neighbor_build is not None:
            line["neighbor_build"] = neighbor_build
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if extra is not None:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        # process-group teardown after captured NCCL collectives can block; every rank is done: leave directly
        barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
