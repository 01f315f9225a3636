"""BASELINE config 5: full-shape inference on a 1M-point synthetic noisy cloud + Chamfer evaluation, phase timings,
with the reference's CPU building blocks (sklearn KDTree radius queries) timed beside on a sample.
usage: python tools/denoise_1m.py [n_points] [num_points]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(dev, n=1_000_000, num_points=8192, cpu=True):
    import bench
    from deep3dpointclouddenoising_b200 import inference, synthetic
    clean = synthetic.make_cloud(0, n, sigma=0.0)
    noisy = (clean + np.random.default_rng(1).standard_normal((n, 3)).astype(np.float32) * 0.005).astype(np.float32)
    dclean, dnoisy = torch.from_numpy(clean).to(dev), torch.from_numpy(noisy).to(dev)
    model, _, cfg = bench.build_model("pospool", num_points)
    model = model.to(dev).eval()

    def timed(fn):
        torch.cuda.synchronize()
        t = time.time()
        out = fn()
        torch.cuda.synchronize()
        return out, (time.time() - t) * 1e3

    inference.patch_centres(dnoisy[:20000].contiguous(), 0.05)  # warm-up of the kernels
    centres, t_centres = timed(lambda: inference.patch_centres(dnoisy, 0.05))
    patches, t_patches = timed(lambda: inference.extract_patches(dnoisy, centres, 0.05, num_points))
    (den, off, votes), t_total = timed(lambda: inference.denoise_cloud(model, dnoisy, 0.05, 0.05, num_points, batch_size=16))
    (ratio, cd_d, cd_n), t_cd = timed(lambda: inference.chamfer_ratio(dclean, dnoisy, den))
    P = centres.shape[0]
    out = {"workload": "configs[4]: full-shape inference (patch centres by grid subsampling at 0.05, radius-0.05 patches, "
                       "batched U-Net forward with random weights, vote averaging) + Chamfer evaluation",
           "n_points": n, "patches": int(P), "num_points": num_points,
           "points_in_ball_mean": float(patches[1].sum(1).float().mean()),
           "ms_patch_centres": round(t_centres, 2), "ms_patch_extraction": round(t_patches, 2),
           "ms_denoise_total": round(t_total, 2), "ms_chamfer_x2": round(t_cd, 2),
           "points_per_s_total": round(n / (t_total / 1e3), 1), "covered_fraction": float((votes > 0).float().mean()),
           "chamfer_ratio_random_weights": float(ratio)}
    if cpu:  # CPU blocks the reference uses, on a bounded sample
        from sklearn.neighbors import KDTree
        t = time.time()
        tree = KDTree(noisy)
        t_build = time.time() - t
        sample = centres[:32].cpu().numpy()
        t = time.time()
        tree.query_radius(noisy[sample], r=0.05, return_distance=True, sort_results=True)
        t_q = (time.time() - t) / 32
        out["cpu_kdtree_build_s"] = round(t_build, 2)
        out["cpu_query_radius_ms_per_patch"] = round(t_q * 1e3, 2)
        out["cpu_patch_extraction_s_extrapolated"] = round(t_build + t_q * P, 1)
    del model, dclean, dnoisy, den, off, votes, patches
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    num_points = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    res = run(torch.device("cuda:0"), n, num_points)
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/denoise_1m.json", "w"), indent=1)
