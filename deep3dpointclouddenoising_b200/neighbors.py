"""Neighbour lists as objects + a per-forward cache.

The reference recomputes the ball query inside every MaskedQueryAndGroup call: 14 queries per forward
for 9 distinct (query set, support set, radius, nsample) combinations (resnet.py:47-68: the strided
bottleneck's max-pool and its local aggregation ask for the same list; la1 and btnk1 too).  Here a
NeighborList is built once per combination and carries everything the fused kernels need:

    idx, idx_mask   (B, M, ns) int32      what _ext.masked_ordered_ball_query returns
    nvalid          (B, M)     int32      in-radius count per query (replaces the dense mask in-kernel)
    rowptr, entries                      inverse map (CSR by support), built lazily for backward

`prebuild` enqueues the whole pyramid of one forward (subsamplings, ball queries, upsampling queries, inverse maps) on a
SIDE stream: these are small-grid, latency-bound integer kernels (2.4 ms back to back for a 16 x 8192 batch) that depend
on coordinates only, so they overlap with the convolutions / BatchNorm / aggregations of the main stream.  Every cached
item carries the CUDA event recorded after its build; the first consumer on another stream waits for it (inside a CUDA
graph capture this becomes a fork / join of two branches).
"""
import torch

from . import ops

_side_streams = {}
_final_events = {}
_building_on_side = [False]  # set while `prebuild` runs: items record a completion event


def _side_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=key)
    return key, _side_streams[key]


class _Produced:
    """Mixin: remembers the stream an item was built on and the event recorded after the build."""
    __slots__ = ()

    def _mark(self):
        self._stream, self._event = None, None
        if _building_on_side[0]:
            self._stream = torch.cuda.current_stream()
            self._event = torch.cuda.Event()
            self._event.record(self._stream)

    def sync(self):
        """Makes the current stream wait for the build if it happened on another stream."""
        ev = self._event
        if ev is not None and torch.cuda.current_stream() != self._stream:
            torch.cuda.current_stream().wait_event(ev)  # the event stays: a consumer on yet another stream waits too
        return self


class NeighborList(_Produced):
    __slots__ = ("idx", "idx_mask", "nvalid", "n_support", "by_support", "_csr", "_keepalive", "_stream", "_event",
                 "_csr_event", "_csr_stream")

    def __init__(self, idx, idx_mask, nvalid, n_support, keepalive=(), by_support=None):
        self.idx, self.idx_mask, self.nvalid, self.n_support = idx, idx_mask, nvalid, n_support
        self.by_support = by_support  # winners in ascending support index (staged-tile kernels), or None
        self._csr = None
        self._csr_event, self._csr_stream = None, None
        self._keepalive = keepalive  # the tensors the cache key points at must outlive the entry
        self._mark()

    @property
    def nsample(self):
        return self.idx.shape[-1] if self.idx.dim() == 3 else 1

    def csr(self):
        if self._csr is None:
            self._csr = ops.build_inverse_map(self.idx, self.n_support)
            if _building_on_side[0]:
                self._csr_stream = torch.cuda.current_stream()
                self._csr_event = torch.cuda.Event()
                self._csr_event.record(self._csr_stream)
        elif self._csr_event is not None and torch.cuda.current_stream() != self._csr_stream:
            torch.cuda.current_stream().wait_event(self._csr_event)
        return self._csr


class _Subsampled(_Produced):
    __slots__ = ("sub_xyz", "sub_mask", "_keepalive", "_stream", "_event")

    def __init__(self, sub_xyz, sub_mask, keepalive):
        self.sub_xyz, self.sub_mask, self._keepalive = sub_xyz, sub_mask, keepalive
        self._mark()


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
            out = ops.ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample, want_nvalid=True,
                                 want_by_support=True)
        idx, msk, nv = out[:3]
        return NeighborList(idx, msk, nv, support_xyz.shape[1], (query_xyz, support_xyz, query_mask, support_mask),
                            by_support=out[3] if len(out) > 3 else None)

    return cache.get("ball", (query_xyz, support_xyz, query_mask, support_mask), (float(radius), int(nsample)), build).sync()


def nearest_neighbors(query_xyz, support_xyz, query_mask, support_mask):
    """Cached masked nearest query -> NeighborList with idx (B, M, 1)."""
    def build():
        with torch.no_grad():
            idx, msk = ops.nearest_query(query_xyz, support_xyz, query_mask, support_mask)
        return NeighborList(idx, msk, None, support_xyz.shape[1], (query_xyz, support_xyz, query_mask, support_mask))

    return cache.get("nearest", (query_xyz, support_xyz, query_mask, support_mask), (), build).sync()


def grid_subsample(xyz, mask, npoint, sample_dl):
    """Cached masked grid subsampling -> (sub_xyz, sub_mask)."""
    def build():
        with torch.no_grad():
            sub_xyz, sub_mask = ops.grid_subsample(xyz, mask, npoint, sample_dl)
        return _Subsampled(sub_xyz, sub_mask, (xyz, mask))

    out = cache.get("grid", (xyz, mask), (int(npoint), float(sample_dl)), build).sync()
    return out.sub_xyz, out.sub_mask


class _Order(_Produced):
    __slots__ = ("order", "_keepalive", "_stream", "_event")

    def __init__(self, order, keepalive):
        self.order, self._keepalive = order, keepalive
        self._mark()


def spatial_order(xyz):
    """Cached Morton processing order (B, N) int32 of a point set for the staged-tile kernels; None when the set is
    larger than they support (ops.TILE_MAX_POINTS) or `runtime.staged_tiles` is off."""
    from .utils.config import runtime
    if not runtime.staged_tiles or xyz.shape[1] > ops.TILE_MAX_POINTS:
        return None

    def build():
        with torch.no_grad():
            return _Order(ops.spatial_order(xyz), (xyz,))

    return cache.get("order", (xyz,), (), build).sync().order


def prebuild(xyz, mask, radius, nsample0, stages, with_csr, with_order=True):
    """Enqueues every neighbourhood structure of one U-Net forward on the side stream and fills the cache.
    stages: per strided stage (sample_dl, npoint, radius_in, nsample_in, radius_out, nsample_out) — the very values the
    modules were constructed with (resnet.py), so their cache keys match."""
    if not (cache.enabled and xyz.is_cuda):
        return
    key, side = _side_stream(xyz.device)
    side.wait_stream(torch.cuda.current_stream())  # fork; also orders the side stream after everything enqueued so far
    _building_on_side[0] = True
    try:
        with torch.cuda.stream(side):
            # in the order the forward pass consumes them (the side stream is in-order: a late item delays its consumer)
            levels, lists = [(xyz, mask)], [ball_neighbors(xyz, xyz, mask, mask, radius, nsample0)]
            if with_order:  # every level runs a self query (PosPool forward as staged tiles)
                spatial_order(xyz)
            for sample_dl, npoint, r_in, ns_in, r_out, ns_out in stages:
                px, pm = levels[-1]
                sx, sm = grid_subsample(px, pm, npoint, sample_dl)
                lists.append(ball_neighbors(sx, px, sm, pm, r_in, ns_in))
                lists.append(ball_neighbors(sx, sx, sm, sm, r_out, ns_out))
                if with_order:
                    spatial_order(sx)
                levels.append((sx, sm))
            ups = [nearest_neighbors(levels[k - 1][0], levels[k][0], levels[k - 1][1], levels[k][1])
                   for k in range(len(levels) - 1, 0, -1)]
            if with_csr:  # in the order backward asks for them: decoder first, then coarse to fine
                for nbr in ups + lists[::-1]:
                    nbr.csr()
            final = torch.cuda.Event()
            final.record(side)
    finally:
        _building_on_side[0] = False
    _final_events[key] = final


def join(device=None):
    """Makes the current stream wait for everything `prebuild` enqueued (required before a graph capture ends)."""
    if not torch.cuda.is_available():
        return
    key = torch.cuda.current_device() if device is None or torch.device(device).index is None else torch.device(device).index
    final = _final_events.pop(key, None)
    if final is not None:
        torch.cuda.current_stream().wait_event(final)
