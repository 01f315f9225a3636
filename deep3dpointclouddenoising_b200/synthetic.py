"""Seeded synthetic "PointCleanNet-shaped" noisy patches.

Shapes and statistics follow what the reference dataset feeds the network
(/root/reference/offset_dataset.py:163-185 diverse noise bins, :630-672 ball patch + padding/mask,
:683,695 pick point to the origin and to index 0; u_net_arch/data_utils.py:198-222 random rotation;
offset_dataset.py:726 features = xyz^T).  There is no dataset on the box, so the surface is a random
quadric height field; everything is numpy on the host (the bench copies it to the GPU per step).
"""
import numpy as np

NOISE_SIGMAS = (0.0, 0.0025, 0.005, 0.01, 0.015, 0.025)


def _rotation(rng):
    ax, ay, az = rng.uniform(-np.pi, np.pi, 3)
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return (rz @ ry @ rx).astype(np.float32)


def make_patch(rng, num_points, in_radius=0.05, valid=None):
    """One patch: (points (N,3) f32, mask (N,) i32, features (3,N) f32, offsets (N,3) f32)."""
    n_valid = num_points if valid is None else int(valid)
    a, b, c = rng.uniform(-8, 8, 3)
    pts = np.empty((0, 3), np.float64)
    while pts.shape[0] < n_valid:
        xy = rng.uniform(-in_radius, in_radius, (2 * n_valid + 64, 2))
        z = a * xy[:, 0] ** 2 + b * xy[:, 0] * xy[:, 1] + c * xy[:, 1] ** 2
        cand = np.concatenate([xy, z[:, None]], 1)
        cand = cand[np.linalg.norm(cand, axis=1) < in_radius]
        pts = np.concatenate([pts, cand], 0)
    clean = pts[:n_valid]
    # diverse noise: six equal bins of sigma, gaussian, clipped to +-0.03, shuffled
    sig = np.repeat(np.array(NOISE_SIGMAS), -(-n_valid // len(NOISE_SIGMAS)))[:n_valid]
    rng.shuffle(sig)
    noise = np.clip(rng.standard_normal((n_valid, 3)) * sig[:, None], -0.03, 0.03)
    noise[0] = 0.0
    clean[0] = 0.0  # the pick point sits at the origin, index 0
    noisy = clean + noise
    rot = _rotation(rng)
    noisy = (noisy @ rot.T).astype(np.float32)
    offs = (-(noise @ rot.T)).astype(np.float32)
    points = np.zeros((num_points, 3), np.float32)
    offsets = np.zeros((num_points, 3), np.float32)
    mask = np.zeros((num_points,), np.int32)
    points[:n_valid], offsets[:n_valid], mask[:n_valid] = noisy, offs, 1
    if n_valid < num_points:  # padding = random duplicates of valid points (offset_dataset.py:660-672)
        dup = rng.integers(0, n_valid, num_points - n_valid)
        points[n_valid:], offsets[n_valid:] = noisy[dup], offs[dup]
    return points, mask, np.ascontiguousarray(points.T), offsets


def make_batch(seed, batch, num_points, in_radius=0.05, ragged=False):
    """Batch of patches.  ragged=False: all-valid masks (throughput mode); True: valid prefix U{0.75N..N}."""
    rng = np.random.default_rng(seed)
    out = [[], [], [], []]
    for _ in range(batch):
        valid = int(rng.integers(int(0.75 * num_points), num_points + 1)) if ragged else None
        for dst, arr in zip(out, make_patch(rng, num_points, in_radius, valid)):
            dst.append(arr)
    return tuple(np.stack(x) for x in out)


def make_cloud(seed, num_points, sigma=0.005):
    """Unit-diameter noisy sphere/torus mix (BASELINE config 1 / 5)."""
    rng = np.random.default_rng(seed)
    half = num_points // 2
    u = rng.standard_normal((half, 3))
    sphere = 0.5 * u / np.linalg.norm(u, axis=1, keepdims=True)
    t, p = rng.uniform(0, 2 * np.pi, (2, num_points - half))
    torus = np.stack([(0.3 + 0.1 * np.cos(p)) * np.cos(t), (0.3 + 0.1 * np.cos(p)) * np.sin(t), 0.1 * np.sin(p)], 1)
    pts = np.concatenate([sphere, torus], 0) + rng.standard_normal((num_points, 3)) * sigma
    rng.shuffle(pts)
    return pts.astype(np.float32)
