"""Neighbour lists as objects + a per-forward cache.

The reference recomputes the ball query inside every MaskedQueryAndGroup call: 14 queries per forward
for 9 distinct (query set, support set, radius, nsample) combinations (resnet.py:47-68: the strided
bottleneck's max-pool and its local aggregation ask for the same list; la1 and btnk1 too).  Here a
NeighborList is built once per combination and carries everything the fused kernels need:

    idx, idx_mask   (B, M, ns) int32      what _ext.masked_ordered_ball_query returns
    nvalid          (B, M)     int32      in-radius count per query (replaces the dense mask in-kernel)
    rowptr, entries                      inverse map (CSR by support), built lazily for backward
"""
import torch

from . import ops


class NeighborList:
    __slots__ = ("idx", "idx_mask", "nvalid", "n_support", "_csr", "_keepalive")

    def __init__(self, idx, idx_mask, nvalid, n_support, keepalive=()):
        self.idx, self.idx_mask, self.nvalid, self.n_support = idx, idx_mask, nvalid, n_support
        self._csr = None
        self._keepalive = keepalive  # the tensors the cache key points at must outlive the entry

    @property
    def nsample(self):
        return self.idx.shape[-1] if self.idx.dim() == 3 else 1

    def csr(self):
        if self._csr is None:
            self._csr = ops.build_inverse_map(self.idx, self.n_support)
        return self._csr


class _Cache:
    """Keyed on tensor identity (data_ptr + version); holds the key tensors alive so that a pointer can
    never be recycled while its entry exists.  Cleared at the start of every backbone forward."""

    def __init__(self, max_entries=64):
        self.entries = {}
        self.max_entries = max_entries
        self.enabled = True
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _tid(t):
        return (t.data_ptr(), t._version, tuple(t.shape))

    def clear(self):
        self.entries.clear()

    def get(self, kind, tensors, scalars, build):
        if not self.enabled:
            return build()
        key = (kind,) + tuple(self._tid(t) for t in tensors) + tuple(scalars)
        hit = self.entries.get(key)
        if hit is not None:
            self.hits += 1
            return hit
        self.misses += 1
        if len(self.entries) >= self.max_entries:
            self.entries.clear()
        val = build()
        self.entries[key] = val
        return val


cache = _Cache()


def ball_neighbors(query_xyz, support_xyz, query_mask, support_mask, radius, nsample):
    """Cached masked ordered ball query -> NeighborList."""
    def build():
        with torch.no_grad():
            idx, msk, nv = ops.ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample,
                                          want_nvalid=True)
        return NeighborList(idx, msk, nv, support_xyz.shape[1], (query_xyz, support_xyz, query_mask, support_mask))

    return cache.get("ball", (query_xyz, support_xyz, query_mask, support_mask), (float(radius), int(nsample)), build)


def nearest_neighbors(query_xyz, support_xyz, query_mask, support_mask):
    """Cached masked nearest query -> NeighborList with idx (B, M, 1)."""
    def build():
        with torch.no_grad():
            idx, msk = ops.nearest_query(query_xyz, support_xyz, query_mask, support_mask)
        return NeighborList(idx, msk, None, support_xyz.shape[1], (query_xyz, support_xyz, query_mask, support_mask))

    return cache.get("nearest", (query_xyz, support_xyz, query_mask, support_mask), (), build)


def grid_subsample(xyz, mask, npoint, sample_dl):
    """Cached masked grid subsampling -> (sub_xyz, sub_mask)."""
    def build():
        with torch.no_grad():
            sub_xyz, sub_mask = ops.grid_subsample(xyz, mask, npoint, sample_dl)
        return sub_xyz, sub_mask, (xyz, mask)

    out = cache.get("grid", (xyz, mask), (int(npoint), float(sample_dl)), build)
    return out[0], out[1]
