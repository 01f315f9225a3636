"""TEST INFRASTRUCTURE — numpy front-ends to the CPU checkers for the index-building ops.

Two families with identical signatures:

* ``restated.*``  -> oracle/liboracle.so, our plain-C restatement (oracle_c.c);
* ``reference.*`` -> oracle/_ref/libref_emul.so, the reference's own CUDA kernel definitions compiled for
  the host by oracle/Makefile (exists only where /root/reference was available at build time, or where
  the prebuilt .so travelled with the repo snapshot).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this
module.  The product package (deep3dpointclouddenoising_b200) never does.
"""
import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_c_int, _c_float = ctypes.c_int, ctypes.c_float


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ref_n_threads(work_size):
    """Launch-shape helper of the reference (include/cuda_utils.h:20-24)."""
    pow_2 = int(math.log(float(work_size)) / math.log(2.0))
    return max(min(1 << pow_2, 512), 1)


def ref_block_config(x, y):
    """include/cuda_utils.h:26-33."""
    xt = ref_n_threads(x)
    yt = max(min(ref_n_threads(y), 512 // xt), 1)
    return xt, yt


class _Restated:
    name = "restated"

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle liboracle.so` (or __graft_entry__.build())")
        L = ctypes.CDLL(path)
        L.oracle_ball_query.argtypes = [_c_int, _c_int, _c_int, _c_float, _c_int, _f32p, _f32p, _i32p, _i32p, _i32p, _i32p]
        L.oracle_nearest_query.argtypes = [_c_int, _c_int, _c_int, _f32p, _f32p, _i32p, _i32p, _i32p, _i32p]
        L.oracle_grid_subsampling.argtypes = [_c_int, _c_int, _c_int, _c_float, _f32p, _i32p, _f32p, _i32p]
        L.oracle_group_points.argtypes = [_c_int] * 5 + [_f32p, _i32p, _f32p]
        L.oracle_group_points_grad.argtypes = [_c_int] * 5 + [_f32p, _i32p, _f32p]
        for f in (L.oracle_ball_query, L.oracle_nearest_query, L.oracle_grid_subsampling, L.oracle_group_points,
                  L.oracle_group_points_grad):
            f.restype = None
        self.L = L

    def ball_query(self, query_xyz, support_xyz, query_mask, support_mask, radius, nsample):
        q, s, qm, sm = _f32(query_xyz), _f32(support_xyz), _i32(query_mask), _i32(support_mask)
        b, m, _ = q.shape
        n = s.shape[1]
        idx = np.zeros((b, m, nsample), np.int32)
        msk = np.zeros((b, m, nsample), np.int32)
        self.L.oracle_ball_query(b, n, m, radius, nsample, q, s, qm, sm, idx, msk)
        return idx, msk

    def nearest_query(self, query_xyz, support_xyz, query_mask, support_mask):
        q, s, qm, sm = _f32(query_xyz), _f32(support_xyz), _i32(query_mask), _i32(support_mask)
        b, m, _ = q.shape
        n = s.shape[1]
        idx = np.zeros((b, m, 1), np.int32)
        msk = np.zeros((b, m, 1), np.int32)
        self.L.oracle_nearest_query(b, n, m, q, s, qm, sm, idx, msk)
        return idx, msk

    def grid_subsampling(self, xyz, mask, npoint, sample_dl):
        p, mk = _f32(xyz), _i32(mask)
        b, n, _ = p.shape
        sub = np.zeros((b, npoint, 3), np.float32)
        subm = np.zeros((b, npoint), np.int32)
        self.L.oracle_grid_subsampling(b, n, npoint, sample_dl, p, mk, sub, subm)
        return sub, subm

    def group_points(self, points, idx):
        p, i = _f32(points), _i32(idx)
        b, c, n = p.shape
        _, npnt, ns = i.shape
        out = np.zeros((b, c, npnt, ns), np.float32)
        self.L.oracle_group_points(b, c, n, npnt, ns, p, i, out)
        return out

    def group_points_grad(self, grad_out, idx, n):
        g, i = _f32(grad_out), _i32(idx)
        b, c, npnt, ns = g.shape
        out = np.zeros((b, c, n), np.float32)
        self.L.oracle_group_points_grad(b, c, n, npnt, ns, g, i, out)
        return out


class _Reference:
    """The reference kernels themselves, executed on the host (oracle/cpu_emul)."""
    name = "reference"

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libref_emul.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: needs /root/reference at build time (`make -C oracle ref`)")
        L = ctypes.CDLL(path)
        L.emul_ball_query.argtypes = [_c_int, _c_int, _c_int, _c_float, _c_int, _f32p, _f32p, _i32p, _i32p, _i32p, _i32p,
                                      _f32p, _i32p, _c_int]
        L.emul_nearest_query.argtypes = [_c_int, _c_int, _c_int, _f32p, _f32p, _i32p, _i32p, _i32p, _i32p, _c_int]
        L.emul_grid_subsampling.argtypes = [_c_int, _c_int, _c_int, _c_float, _f32p, _i32p, _f32p, _i32p, _i32p, _i32p, _f32p]
        L.emul_group_points.argtypes = [_c_int] * 5 + [_f32p, _i32p, _f32p, _c_int, _c_int]
        L.emul_group_points_grad.argtypes = [_c_int] * 5 + [_f32p, _i32p, _f32p, _c_int, _c_int]
        for f in (L.emul_ball_query, L.emul_nearest_query, L.emul_grid_subsampling, L.emul_group_points,
                  L.emul_group_points_grad):
            f.restype = None
        self.L = L

    # scratch tensors are zero-filled like masked_ordered_ball_query.cpp:38-44 / masked_grid_subsampling.cpp:19-34
    def ball_query(self, query_xyz, support_xyz, query_mask, support_mask, radius, nsample):
        q, s, qm, sm = _f32(query_xyz), _f32(support_xyz), _i32(query_mask), _i32(support_mask)
        b, m, _ = q.shape
        n = s.shape[1]
        idx = np.zeros((b, m, nsample), np.int32)
        msk = np.zeros((b, m, nsample), np.int32)
        dists = np.zeros((b, m, 3 * nsample), np.float32)
        tmp = np.zeros((b, m, 3 * nsample), np.int32)
        self.L.emul_ball_query(b, n, m, radius, nsample, q, s, qm, sm, idx, msk, dists, tmp, ref_n_threads(m))
        return idx, msk

    def nearest_query(self, query_xyz, support_xyz, query_mask, support_mask):
        q, s, qm, sm = _f32(query_xyz), _f32(support_xyz), _i32(query_mask), _i32(support_mask)
        b, m, _ = q.shape
        n = s.shape[1]
        idx = np.zeros((b, m, 1), np.int32)
        msk = np.zeros((b, m, 1), np.int32)
        self.L.emul_nearest_query(b, n, m, q, s, qm, sm, idx, msk, ref_n_threads(m))
        return idx, msk

    def grid_subsampling(self, xyz, mask, npoint, sample_dl):
        p, mk = _f32(xyz), _i32(mask)
        b, n, _ = p.shape
        sub = np.zeros((b, npoint, 3), np.float32)
        subm = np.zeros((b, npoint), np.int32)
        mapidx = np.zeros((b, n), np.int32)
        tmpidx = np.zeros((b, n), np.int32)
        tmpxyz = np.zeros((b, n, 3), np.float32)
        self.L.emul_grid_subsampling(b, n, npoint, sample_dl, p, mk, sub, subm, mapidx, tmpidx, tmpxyz)
        return sub, subm

    def group_points(self, points, idx):
        p, i = _f32(points), _i32(idx)
        b, c, n = p.shape
        _, npnt, ns = i.shape
        out = np.zeros((b, c, npnt, ns), np.float32)
        tx, ty = ref_block_config(npnt, c)
        self.L.emul_group_points(b, c, n, npnt, ns, p, i, out, tx, ty)
        return out

    def group_points_grad(self, grad_out, idx, n):
        g, i = _f32(grad_out), _i32(idx)
        b, c, npnt, ns = g.shape
        out = np.zeros((b, c, n), np.float32)
        tx, ty = ref_block_config(npnt, c)
        self.L.emul_group_points_grad(b, c, n, npnt, ns, g, i, out, tx, ty)
        return out


class _RefGridSubCPU:
    """Reference CPU grid subsampling (cpp_wrappers/cpp_subsampling) behind oracle/gridsub_shim.cpp."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libref_gridsub.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = ctypes.CDLL(path)
        L.ref_grid_subsampling.argtypes = [_f32p, _c_int, ctypes.c_void_p, _c_int, ctypes.c_void_p, _c_int, _c_float,
                                           _f32p, ctypes.c_void_p, ctypes.c_void_p]
        L.ref_grid_subsampling.restype = _c_int
        self.L = L

    def compute(self, points, sample_dl):
        p = _f32(points)
        out = np.zeros_like(p)
        k = self.L.ref_grid_subsampling(p, p.shape[0], None, 0, None, 0, sample_dl, out, None, None)
        return out[:k].copy()


class _RefNanoflann:
    """oracle/_ref/libref_nanoflann.so: radius search over the reference's vendored nanoflann.hpp and its PointCloud
    adaptor (cpp_wrappers/cpp_utils/nanoflann/nanoflann.hpp:1280, cpp_utils/cloud/cloud.h:151-175)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libref_nanoflann.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing (built by `make -C oracle ref` where /root/reference exists)")
        L = ctypes.CDLL(path)
        L.ref_nanoflann_radius.argtypes = [_f32p, _c_int, _f32p, _c_int, _c_float, _c_int, _c_int, _i32p, _i32p]
        L.ref_nanoflann_radius.restype = _c_int
        self.L = L

    def radius(self, support, query, radius, cap, threads=1):
        """-> (idx (m, cap) int32 nearest-first, -1 padded; count (m,) neighbours inside the ball)."""
        s, q = _f32(support), _f32(query)
        idx = np.empty((q.shape[0], cap), np.int32)
        cnt = np.empty((q.shape[0],), np.int32)
        self.L.ref_nanoflann_radius(s, s.shape[0], q, q.shape[0], radius, cap, threads, idx, cnt)
        return idx, cnt


_cache = {}


def restated():
    if "restated" not in _cache:
        _cache["restated"] = _Restated()
    return _cache["restated"]


def reference():
    if "reference" not in _cache:
        _cache["reference"] = _Reference()
    return _cache["reference"]


def reference_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_emul.so"))


def ref_gridsub_cpu():
    if "gridsub" not in _cache:
        _cache["gridsub"] = _RefGridSubCPU()
    return _cache["gridsub"]


def ref_nanoflann():
    if "nanoflann" not in _cache:
        _cache["nanoflann"] = _RefNanoflann()
    return _cache["nanoflann"]
