"""BASELINE config 5: full-shape inference on a 1M-point synthetic noisy cloud + Chamfer evaluation, phase timings,
with the reference's CPU building blocks (sklearn KDTree radius queries, cKDTree Chamfer) timed beside on a sample."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deep3dpointclouddenoising_b200 import inference, ops, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
num_points = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
clean = synthetic.make_cloud(0, n, sigma=0.0)
noisy = (clean + np.random.default_rng(1).standard_normal((n, 3)).astype(np.float32) * 0.005).astype(np.float32)
dclean, dnoisy = torch.from_numpy(clean).to(dev), torch.from_numpy(noisy).to(dev)
model, _, cfg = bench.build_model("pospool", num_points)
model = model.to(dev).eval()

def timed(fn):
    torch.cuda.synchronize(); t = time.time(); out = fn(); torch.cuda.synchronize(); return out, (time.time() - t) * 1e3

inference.patch_centres(dnoisy[:20000].contiguous(), 0.05)  # warm-up of the kernels
centres, t_centres = timed(lambda: inference.patch_centres(dnoisy, 0.05))
patches, t_patches = timed(lambda: inference.extract_patches(dnoisy, centres, 0.05, num_points))
(den, off, votes), t_total = timed(lambda: inference.denoise_cloud(model, dnoisy, 0.05, 0.05, num_points, batch_size=16))
(ratio, cd_d, cd_n), t_cd = timed(lambda: inference.chamfer_ratio(dclean, dnoisy, den))
P = centres.shape[0]
out = {"n_points": n, "patches": int(P), "num_points": num_points, "points_in_ball_mean": float(patches[1].sum(1).float().mean()),
       "ms_patch_centres": t_centres, "ms_patch_extraction": t_patches, "ms_denoise_total": t_total, "ms_chamfer_x2": t_cd,
       "points_per_s_total": n / (t_total / 1e3), "covered_fraction": float((votes > 0).float().mean()),
       "chamfer_ratio_random_weights": ratio}
# CPU blocks the reference uses, on a bounded sample
from sklearn.neighbors import KDTree
t = time.time(); tree = KDTree(noisy); t_build = time.time() - t
sample = centres[:64].cpu().numpy()
t = time.time(); tree.query_radius(noisy[sample], r=0.05, return_distance=True, sort_results=True); t_q = (time.time() - t) / 64
out["cpu_kdtree_build_s"] = t_build
out["cpu_query_radius_ms_per_patch"] = t_q * 1e3
out["cpu_patch_extraction_s_extrapolated"] = t_build + t_q * P
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/denoise_1m.json", "w"), indent=1)
