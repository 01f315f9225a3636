"""BASELINE configs[0]: the reference's CPU preprocessing — cpp_wrappers grid subsampling + batch radius neighbours over
the 5 U-Net levels — on one synthetic 100k-point noisy shape (unit diameter, sigma 0.005, seed 0), at the geometry
BASELINE.md §3 states: level l = 1..4 subsamples at dl = 0.0015625 * 2^l; neighbour lists are the U-Net's nine:
level-0 self neighbours (r = 0.025), and per level l the strided list (level l queries into level l-1, r = 0.025 * 2^(l-1))
and the self list (r = 0.025 * 2^l), each truncated to the level's nsample nearest ([52, 39, 32, 26, 26]).

GPU leg: d3d_voxel_ids / d3d_voxel_barycentres (grid subsampling, bit-identical to the reference's C++ as sets) and
d3d_radius_patches (nearest-first radius lists).  Host legs, beside it: the reference's own grid_subsampling.cpp
(oracle/_ref shim, single thread like the reference), radius search over the reference's vendored nanoflann.hpp +
PointCloud adaptor (oracle/_ref/libref_nanoflann.so; 1 thread and all cores) and sklearn's KDTree.query_radius (the class
the reference dataset uses, offset_dataset.py:630).  Parity of the legs: tests/test_gpu_inference.py.
usage: python tools/pyramid_100k.py [out.json]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3dpointclouddenoising_b200 import inference, ops, synthetic  # noqa: E402

BASE_DL, BASE_RADIUS, NSAMPLES = 0.0015625, 0.025, [52, 39, 32, 26, 26]


def list_specs():
    """(query level, support level, radius, cap) of the nine neighbour lists."""
    specs = [(0, 0, BASE_RADIUS, NSAMPLES[0])]
    for l in range(1, 5):
        specs.append((l, l - 1, BASE_RADIUS * 2 ** (l - 1), NSAMPLES[l - 1]))
        specs.append((l, l, BASE_RADIUS * 2 ** l, NSAMPLES[l]))
    return specs


def radius_lists(support, query, radius, cap):
    return ops.radius_neighbors(support, query, radius, cap)  # candidate storage tier chosen from the data


def gpu_pyramid(pts):
    levels = [pts]
    for l in range(1, 5):
        sub, _ = inference.voxel_barycentres(levels[-1], BASE_DL * 2 ** l)
        levels.append(sub.contiguous())
    lists = [radius_lists(levels[ls], levels[lq], r, cap) for lq, ls, r, cap in list_specs()]
    return levels, lists


def run(dev, n_rep=5, cpu=True, single_thread=True):
    pts_np = synthetic.make_cloud(0, 100_000, sigma=0.005)
    pts = torch.from_numpy(pts_np).to(dev)
    gpu_pyramid(pts)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n_rep):
        levels, lists = gpu_pyramid(pts)
    t1.record()
    torch.cuda.synchronize()
    gpu_ms = t0.elapsed_time(t1) / n_rep
    sizes = [int(lv.shape[0]) for lv in levels]
    n_queries = sum(sizes[lq] for lq, _, _, _ in list_specs())
    line = {"workload": "configs[0]: 4 grid subsamplings (dl = 0.0015625 * 2^l) + the U-Net's 9 radius neighbour lists "
                        "(r = 0.025 * 2^(l-1) strided, 0.025 * 2^l self; nsample nearest) of one 100k-point noisy cloud",
            "level_points": sizes, "query_points": n_queries, "gpu_ms": round(gpu_ms, 3),
            "gpu_mpts_per_s": round(n_queries / gpu_ms / 1e3, 2),
            "mean_in_radius": [round(float(c.float().mean()), 1) for _, c in lists]}
    if not cpu:
        return line
    from oracle import cpu_index_ops
    host = {"cores": os.cpu_count()}
    subs = [pts_np]
    try:
        ref = cpu_index_ops.ref_gridsub_cpu()
        t = time.time()
        for l in range(1, 5):
            subs.append(ref.compute(subs[-1], BASE_DL * 2 ** l))
        host["grid_subsampling_cpp_ms"] = round((time.time() - t) * 1e3, 1)
    except (FileNotFoundError, OSError) as e:
        subs = [lv.cpu().numpy() for lv in levels]
        host["grid_subsampling_cpp_ms"] = None
        host["grid_subsampling_cpp"] = f"unavailable: {e}"
    try:
        nf = cpu_index_ops.ref_nanoflann()
        legs = ((1, "nanoflann_radius_1thread_ms"),) if single_thread else ()
        for threads, key in legs + ((os.cpu_count() or 1, "nanoflann_radius_all_cores_ms"),):
            t = time.time()
            for lq, ls, r, cap in list_specs():
                nf.radius(subs[ls], subs[lq], r, cap, threads)
            host[key] = round((time.time() - t) * 1e3, 1)
    except (FileNotFoundError, OSError) as e:
        host["nanoflann"] = f"unavailable: {e}"
    from sklearn.neighbors import KDTree
    t = time.time()
    for lq, ls, r, cap in list_specs()[:3]:  # the three finest lists only: sklearn materialises every neighbour
        KDTree(subs[ls]).query_radius(subs[lq], r=r)
    host["sklearn_kdtree_radius_first3_lists_ms"] = round((time.time() - t) * 1e3, 1)
    line["host"] = host
    best = host.get("nanoflann_radius_all_cores_ms")
    if best is not None and host.get("grid_subsampling_cpp_ms") is not None:
        line["host_total_ms"] = round(best + host["grid_subsampling_cpp_ms"], 1)
    return line


if __name__ == "__main__":
    out = run(torch.device("cuda:0"))
    print(json.dumps(out))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
