"""Kernel points and weight initialisation for PseudoGrid (the file name keeps the reference's spelling,
u_net_arch/models/utlis.py).

    create_kernel_points(radius, num_kpoints, num_kernels, dimension, fixed)   ref utlis.py:153-284
    radius_gaussian(sq_r, sig, eps)                                            ref utlis.py:287-294
    weight_variable(size)                                                      ref utlis.py:297-303

Lookup order for a disposition (same file naming as the reference, so existing kernel folders work):
  1. $JOB_LOAD_DIR/kernels/dispositions, else $JOB_LOG_DIR/kernels/dispositions   (ref :158-165)
  2. the table shipped with the reference (models/kernel_dispositions.py): the five radii 0.015*2^l that
     the train_dist.py geometry produces, bit-identical to the reference's .npy fixtures
  3. otherwise: unit disposition (table, or a fresh repulsion optimisation) scaled by `radius`, rotated
     by a random orthonormal frame, plus 1 % noise — like the reference's generator (:222-275) —
     and saved to the kernel folder when one is configured.
"""
import os

import numpy as np
import torch

from . import kernel_dispositions as _table


def _kernel_dir():
    load_dir, log_dir = os.environ.get("JOB_LOAD_DIR"), os.environ.get("JOB_LOG_DIR")
    base = load_dir if load_dir is not None else log_dir
    return os.path.join(base, "kernels", "dispositions") if base is not None else None


def _optimise_unit_disposition(num_kpoints, dimension, fixed, rng, iters=3000):
    """Points in the unit ball pushed apart by 1/d^2 repulsion with a pull towards the centre; 'center'
    pins point 0 at the origin, 'verticals' additionally pins points 1, 2 on the z axis."""
    pts = rng.uniform(-1, 1, (num_kpoints * 4, dimension))
    pts = pts[np.linalg.norm(pts, axis=1) < 0.9][:num_kpoints]
    while pts.shape[0] < num_kpoints:
        pts = np.concatenate([pts, rng.uniform(-0.5, 0.5, (1, dimension))], 0)
    if fixed in ("center", "verticals"):
        pts[0] = 0
    if fixed == "verticals" and num_kpoints >= 3:
        pts[1], pts[2] = 0, 0
        pts[1, -1], pts[2, -1] = 0.66, -0.66
    step = 0.01
    for _ in range(iters):
        diff = pts[:, None, :] - pts[None, :, :]
        d2 = (diff ** 2).sum(-1) + 1e-6
        np.fill_diagonal(d2, np.inf)
        grad = (diff / d2[..., None] ** 1.5).sum(1) * 0.05 - 2 * pts
        if fixed in ("center", "verticals"):
            grad[0] = 0
        if fixed == "verticals":
            grad[1:3, :-1] = 0
        pts = pts + step * grad / max(np.abs(grad).max(), 1e-9) * 0.1
        step *= 0.9995
    scale = np.linalg.norm(pts, axis=1).max()
    return pts / max(scale, 1e-9)


def create_kernel_points(radius, num_kpoints, num_kernels, dimension, fixed):
    name = "sk_pt_{:04f}_{:03d}_{:s}{}.npy".format(radius, num_kpoints, fixed, "" if dimension == 3 else "_2D")
    folder = _kernel_dir()
    if folder is not None and os.path.exists(os.path.join(folder, name)):
        return np.load(os.path.join(folder, name))
    if dimension == 3 and num_kernels == 1:
        shipped = _table.SCALED.get(("{:04f}".format(radius), num_kpoints, fixed))
        if shipped is not None:
            return np.asarray(shipped, dtype=np.float64).reshape(1, num_kpoints, 3)
    rng = np.random.default_rng(int(round(radius * 1e6)) + 1000 * num_kpoints)
    unit = _table.UNIT.get((num_kpoints, fixed)) if dimension == 3 else None
    unit = np.asarray(unit, np.float64) if unit is not None else _optimise_unit_disposition(num_kpoints, dimension, fixed, rng)
    if dimension == 2:
        return unit
    frames = []
    for _ in range(num_kernels):
        if fixed == "verticals":
            t = rng.uniform(0, 2 * np.pi)
            frames.append(np.array([[np.cos(t), np.sin(t), 0], [-np.sin(t), np.cos(t), 0], [0, 0, 1]]))
        else:
            q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
            frames.append(q)
    kernels = radius * unit[None] @ np.stack(frames)
    if fixed != "verticals":
        kernels = kernels + rng.normal(scale=radius * 0.01, size=kernels.shape)
    if folder is not None:
        try:
            os.makedirs(folder, exist_ok=True)
            np.save(os.path.join(folder, name), kernels)
            open(os.path.join(folder, name + "_finish"), "w").write("finish!")
        except OSError:
            pass
    return kernels


def radius_gaussian(sq_r, sig, eps=1e-9):
    sig = torch.as_tensor(sig, dtype=sq_r.dtype, device=sq_r.device)
    return torch.exp(-sq_r / (2 * sig ** 2 + eps))


def weight_variable(size):
    """Truncated-normal init, std sqrt(2 / fan) with values beyond 2 std zeroed (ref utlis.py:297-303);
    drawn from numpy's global RNG like the reference so that np.random.seed reproduces it."""
    std = np.sqrt(2 / size[-1])
    w = np.random.normal(scale=std, size=size)
    w[np.abs(w) > 2 * std] = 0
    return torch.nn.Parameter(torch.from_numpy(w).float(), requires_grad=True)
