"""TEST INFRASTRUCTURE — ctypes front-end to oracle/_ref/libref_cuda.so: the reference's own CUDA kernels compiled
by nvcc for sm_100a (oracle/Makefile, sources read from /root/reference at build time, never copied), behind the
launch shim oracle/cuda_ref/ref_launch.cu.  This is the arbiter for bit-exactness on the GPU box and the
"kernel to beat" for timings.  Scratch tensors are zero-filled like the reference's torch::zeros."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_cuda.so")


def available():
    return os.path.exists(PATH) and torch.cuda.is_available()


class RefCuda:
    def __init__(self):
        L = ctypes.CDLL(PATH)
        vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.ref_cuda_ball_query.argtypes = [i, i, i, f, i] + [vp] * 8 + [vp]
        L.ref_cuda_grid_subsampling.argtypes = [i, i, i, f] + [vp] * 7 + [vp]
        L.ref_cuda_nearest_query.argtypes = [i, i, i] + [vp] * 6 + [vp]
        L.ref_cuda_group_points.argtypes = [i] * 5 + [vp] * 3 + [vp]
        L.ref_cuda_group_points_grad.argtypes = [i] * 5 + [vp] * 3 + [vp]
        L.ref_cuda_set_malloc_heap.argtypes = [ctypes.c_size_t]
        rc = L.ref_cuda_set_malloc_heap(2 << 30)  # see ref_launch.cu: the reference needs > 8 MB of device heap
        assert rc == 0, rc
        self.L = L

    @staticmethod
    def _st():
        return torch.cuda.current_stream().cuda_stream

    def ball_query(self, q, s, qm, sm, radius, ns):
        B, M, N = q.shape[0], q.shape[1], s.shape[1]
        idx = torch.zeros((B, M, ns), dtype=torch.int32, device=q.device)
        msk = torch.zeros_like(idx)
        dists = torch.zeros((B, M, 3 * ns), dtype=torch.float32, device=q.device)
        tmp = torch.zeros((B, M, 3 * ns), dtype=torch.int32, device=q.device)
        rc = self.L.ref_cuda_ball_query(B, N, M, radius, ns, q.data_ptr(), s.data_ptr(), qm.data_ptr(), sm.data_ptr(),
                                        idx.data_ptr(), msk.data_ptr(), dists.data_ptr(), tmp.data_ptr(), self._st())
        assert rc == 0, rc
        return idx, msk

    def grid_subsampling(self, xyz, mask, m, dl):
        B, N = xyz.shape[0], xyz.shape[1]
        sub = torch.zeros((B, m, 3), dtype=torch.float32, device=xyz.device)
        subm = torch.zeros((B, m), dtype=torch.int32, device=xyz.device)
        a = torch.zeros((B, N), dtype=torch.int32, device=xyz.device)
        b = torch.zeros((B, N), dtype=torch.int32, device=xyz.device)
        c = torch.zeros((B, N, 3), dtype=torch.float32, device=xyz.device)
        rc = self.L.ref_cuda_grid_subsampling(B, N, m, dl, xyz.data_ptr(), mask.data_ptr(), sub.data_ptr(), subm.data_ptr(),
                                              a.data_ptr(), b.data_ptr(), c.data_ptr(), self._st())
        assert rc == 0, rc
        return sub, subm

    def nearest_query(self, q, s, qm, sm):
        B, M, N = q.shape[0], q.shape[1], s.shape[1]
        idx = torch.zeros((B, M, 1), dtype=torch.int32, device=q.device)
        msk = torch.zeros_like(idx)
        rc = self.L.ref_cuda_nearest_query(B, N, M, q.data_ptr(), s.data_ptr(), qm.data_ptr(), sm.data_ptr(),
                                           idx.data_ptr(), msk.data_ptr(), self._st())
        assert rc == 0, rc
        return idx, msk

    def group_points(self, points, idx):
        B, C, N = points.shape
        M, ns = idx.shape[1], idx.shape[2]
        out = torch.zeros((B, C, M, ns), dtype=torch.float32, device=points.device)
        rc = self.L.ref_cuda_group_points(B, C, N, M, ns, points.data_ptr(), idx.data_ptr(), out.data_ptr(), self._st())
        assert rc == 0, rc
        return out

    def group_points_grad(self, grad_out, idx, n):
        B, C, M, ns = grad_out.shape
        out = torch.zeros((B, C, n), dtype=torch.float32, device=grad_out.device)
        rc = self.L.ref_cuda_group_points_grad(B, C, n, M, ns, grad_out.data_ptr(), idx.data_ptr(), out.data_ptr(), self._st())
        assert rc == 0, rc
        return out
