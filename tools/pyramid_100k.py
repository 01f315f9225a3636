"""BASELINE configs[0]: the reference's CPU preprocessing — cpp_wrappers grid_subsampling + batch radius neighbours over
5 U-Net levels — on one synthetic 100k-point noisy shape, on the GPU path (d3d_voxel_ids / d3d_voxel_barycentres /
d3d_radius_patches) and, beside it, on the host: the reference's own grid_subsampling.cpp (oracle/_ref shim, single
thread like the reference) and sklearn's KDTree.query_radius (the class the reference dataset uses,
offset_dataset.py:630).  Parity of both legs is checked in tests/test_gpu_inference.py.
usage: python tools/pyramid_100k.py [out.json]"""
import json
import os
import sys
import time

import torch
from sklearn.neighbors import KDTree

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import inference, ops, synthetic  # noqa: E402

LEVELS, DL0, RADIUS_FACTOR, CAP = 5, 0.01, 2.5, 64


def gpu_pyramid(pts):
    out, cur = [], pts
    for lv in range(LEVELS):
        dl = DL0 * 2 ** lv
        sub, _ = inference.voxel_barycentres(cur, dl)
        idx, cnt = ops.radius_patches(sub, sub.contiguous(), RADIUS_FACTOR * dl, CAP)
        out.append((sub, idx, cnt))
        cur = sub
    return out


def main():
    dev = torch.device("cuda:0")
    pts_np = synthetic.make_cloud(0, 100_000, sigma=0.005)
    pts = torch.from_numpy(pts_np).to(dev)
    gpu_pyramid(pts)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    n_rep = 10
    for _ in range(n_rep):
        levels = gpu_pyramid(pts)
    t1.record()
    torch.cuda.synchronize()
    gpu_ms = t0.elapsed_time(t1) / n_rep
    sizes = [int(lv[0].shape[0]) for lv in levels]

    # host: reference C++ grid subsampling (single thread) + KDTree radius neighbours (all levels)
    cpu = {"grid_subsampling_cpp_ms": None, "kdtree_radius_ms": None}
    cur = pts_np
    try:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import cpu_index_ops
        ref = cpu_index_ops.ref_gridsub_cpu()
        t = time.time()
        subs = []
        for lv in range(LEVELS):
            cur = ref.compute(cur, DL0 * 2 ** lv)
            subs.append(cur)
        cpu["grid_subsampling_cpp_ms"] = (time.time() - t) * 1e3
    except (FileNotFoundError, OSError, ImportError) as e:  # the shim did not travel: time KDTree on our own levels
        subs = [lv[0].cpu().numpy() for lv in levels]
        cpu["grid_subsampling_cpp"] = f"unavailable: {e}"
    t = time.time()
    for lv, sub in enumerate(subs):
        KDTree(sub).query_radius(sub, r=RADIUS_FACTOR * DL0 * 2 ** lv)
    cpu["kdtree_radius_ms"] = (time.time() - t) * 1e3
    line = {"workload": "configs[0]: 5-level grid subsampling + radius neighbours, one 100k-point noisy cloud",
            "level_points": sizes, "dl0": DL0, "radius": f"{RADIUS_FACTOR} x dl", "neighbour_cap": CAP,
            "gpu_ms": round(gpu_ms, 3), "gpu_mpts_per_s": round(100_000 / gpu_ms / 1e3, 2),
            "cpu": cpu, "cpu_cores": os.cpu_count(),
            "cpu_total_ms": None if cpu["grid_subsampling_cpp_ms"] is None else round(cpu["grid_subsampling_cpp_ms"] + cpu["kdtree_radius_ms"], 1)}
    print(json.dumps(line))
    if len(sys.argv) > 1:
        json.dump(line, open(sys.argv[1], "w"))


if __name__ == "__main__":
    main()
