"""Neighbour lists as objects + a per-forward cache.

The reference recomputes the ball query inside every MaskedQueryAndGroup call: 14 queries per forward
for 9 distinct (query set, support set, radius, nsample) combinations (resnet.py:47-68: the strided
bottleneck's max-pool and its local aggregation ask for the same list; la1 and btnk1 too).  Here a
NeighborList is built once per combination and carries everything the fused kernels need:

    idx, idx_mask   (B, M, ns) int32      what _ext.masked_ordered_ball_query returns
    nvalid          (B, M)     int32      in-radius count per query (replaces the dense mask in-kernel)
    rowptr, entries                      inverse map (CSR by support), built lazily for backward
    by_support      (B, M, ns) int32      the winners in ascending support index, (distance rank << 16) | index
    tile plan       uint8 buffer          per tile of 128 Morton-adjacent queries: union of the gathered support rows, the
                                          rank of every list entry in it, and the inverse view (staged-tile PosPool kernels)

`prebuild` enqueues the whole pyramid of one forward (subsamplings, ball queries, upsampling queries, inverse maps) on a
SIDE stream: these are small-grid, latency-bound integer kernels (2.4 ms back to back for a 16 x 8192 batch) that depend
on coordinates only, so they overlap with the convolutions / BatchNorm / aggregations of the main stream.  Every cached
item carries the CUDA event recorded after its build; the first consumer on another stream waits for it (inside a CUDA
graph capture this becomes a fork / join of two branches).  `prefetch` / `adopt` / `fold_pending_into_current` move the
build of the NEXT batch's pyramid into the current step (one-step pipeline, see the section at the end of the file).
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
                 "_csr_event", "_csr_stream", "_plan", "_plan_event", "_plan_stream")

    def __init__(self, idx, idx_mask, nvalid, n_support, keepalive=(), by_support=None):
        self.idx, self.idx_mask, self.nvalid, self.n_support = idx, idx_mask, nvalid, n_support
        self.by_support = by_support  # winners in ascending support index (staged-tile kernels), or None
        self._csr = None
        self._csr_event, self._csr_stream = None, None
        self._plan, self._plan_event, self._plan_stream = None, None, None
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


    def tile_plan(self, query_xyz, query_mask):
        """Tile plan of this list under the Morton order of its queries (staged-tile PosPool kernels), built once; None
        when the list has no by-support form or the sizes are beyond the kernels' limits."""
        if self._plan is None:
            order = spatial_order(query_xyz)
            if order is None or self.by_support is None or self.idx.shape[-1] > ops.TILE_MAX_NSAMPLE \
                    or self.n_support > ops.TILE_MAX_POINTS:
                return None
            with torch.no_grad():
                self._plan = ops.tile_plan(self.by_support, self.nvalid, query_mask, order, self.n_support)
            if _building_on_side[0]:
                self._plan_stream = torch.cuda.current_stream()
                self._plan_event = torch.cuda.Event()
                self._plan_event.record(self._plan_stream)
        elif self._plan_event is not None and torch.cuda.current_stream() != self._plan_stream:
            torch.cuda.current_stream().wait_event(self._plan_event)
        return self._plan


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
        self.arena = None  # the ops.Arena the entries were carved from (prefetched pyramids only)
        self.max_entries = max_entries
        self.enabled = True
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _tid(t):
        return (t.data_ptr(), t._version, tuple(t.shape))

    def clear(self):
        self.entries = {}  # a new dict: an adopted pyramid's dict may still be referenced by its producer
        self.arena = None

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
                lists[0].tile_plan(xyz, mask)
            for sample_dl, npoint, r_in, ns_in, r_out, ns_out in stages:
                px, pm = levels[-1]
                sx, sm = grid_subsample(px, pm, npoint, sample_dl)
                lists.append(ball_neighbors(sx, px, sm, pm, r_in, ns_in))
                lists.append(ball_neighbors(sx, sx, sm, sm, r_out, ns_out))
                if with_order:
                    spatial_order(sx)
                    lists[-1].tile_plan(sx, sm)
                    lists[-2].tile_plan(sx, sm)  # strided list: the scatter-form backward, and the forward on coarse levels
                levels.append((sx, sm))
            ups = [nearest_neighbors(levels[k - 1][0], levels[k][0], levels[k - 1][1], levels[k][1])
                   for k in range(len(levels) - 1, 0, -1)]
            if with_csr:  # in the order backward asks for them: decoder first, then coarse to fine
                from .utils.config import runtime
                # self lists feed local aggregation only: PosPool's scatter-form backward works from the tile plan
                skip_self = with_order and runtime.staged_tiles and runtime.staged_tiles_backward in ('scatter', 'ordered')
                for k, nbr in [(None, u) for u in ups] + list(enumerate(lists))[::-1]:
                    if not (skip_self and k is not None and k % 2 == 0 and nbr._plan is not None):
                        nbr.csr()
            final = torch.cuda.Event()
            final.record(side)
    finally:
        _building_on_side[0] = False
    _final_events[key] = final


# ------------------------------------------------------------------------------------------------
# cross-step pipelining: the pyramid of the NEXT batch, built while this batch's backward runs
# ------------------------------------------------------------------------------------------------
_pending = {}     # device index -> (source key, entries, final event, keepalive, arena): adopted by the next forward on it
_persistent = []  # pyramids a captured graph reads and overwrites at fixed addresses: never freed
_arena_bytes = {}  # (shapes, geometry) -> bytes of one pyramid's outputs, measured by the first build


def _source_key(xyz, mask):
    return (cache._tid(xyz), cache._tid(mask))


def prefetch(xyz, mask, radius, nsample0, stages, with_csr=True, with_order=True):
    """Builds the whole pyramid of (xyz, mask) NOW on the side stream, into a pending cache that the next backbone forward
    on these very tensors adopts instead of building its own (ResNet.prefetch_neighbors is the caller with the module's
    geometry).  Called between a step's forward and its backward, the latency-bound neighbourhood kernels of batch k+1
    (the reference rebuilds them inside every forward, resnet.py:47-68) overlap the backward pass of batch k — the
    pyramid depends on coordinates only, never on the weights, so the training arithmetic is unchanged."""
    if not (cache.enabled and xyz.is_cuda):
        return
    key, side = _side_stream(xyz.device)
    # every output of the build comes out of ONE block (ops.Arena): the graph form of the hand-over is then one copy.
    # The block's size follows from the shapes and the geometry; the first build of a kind measures it (and is redone).
    spec = (tuple(xyz.shape), float(radius), int(nsample0), tuple(stages), bool(with_csr), bool(with_order))
    for attempt in range(2):
        nbytes = _arena_bytes.get(spec, 0)
        side.wait_stream(torch.cuda.current_stream())  # fork first: under a graph capture the block must come from the graph's pool
        with torch.cuda.stream(side):  # the block belongs to the side stream's pool, like the tensors of a plain build
            arena = ops.Arena(nbytes, xyz.device)
        saved, cache.entries = cache.entries, {}
        try:
            with arena:
                prebuild(xyz, mask, radius, nsample0, stages, with_csr, with_order)
            built = cache.entries
        finally:
            cache.entries = saved
        if nbytes:
            break
        _arena_bytes[spec] = arena.need
    _pending[key] = (_source_key(xyz, mask), built, _final_events.pop(key, None), (xyz, mask), arena)


def adopt(xyz, mask):
    """True if a prefetched pyramid of exactly these tensors was pending: it becomes the forward's cache."""
    if not xyz.is_cuda:
        return False
    key, _ = _side_stream(xyz.device)
    pending = _pending.pop(key, None)
    if pending is None or pending[0] != _source_key(xyz, mask):
        return False
    cache.entries, cache.arena = pending[1], pending[4]
    if pending[2] is not None:
        _final_events[key] = pending[2]  # the forward's join() waits for the build as a whole
    return True


def _tensors(item):
    if isinstance(item, NeighborList):
        csr = item._csr if item._csr is not None else (None, None)
        return [item.idx, item.idx_mask, item.nvalid, item.by_support, csr[0], csr[1], item._plan]
    if isinstance(item, _Subsampled):
        return [item.sub_xyz, item.sub_mask]
    return [item.order]


def settle(device=None):
    """Waits (host side) for the pending pyramid and drops its events: what a CUDA-graph capture needs of a pyramid that
    was built before the capture began (a captured stream must not wait on events recorded outside the capture)."""
    key, _ = _side_stream(torch.cuda.current_device() if device is None else device)
    pending = _pending.get(key)
    if pending is None:
        return
    torch.cuda.synchronize(key)
    for item in pending[1].values():
        item._event = None
        if isinstance(item, NeighborList):
            item._csr_event = None
            item._plan_event = None
    _pending[key] = (pending[0], pending[1], None, pending[3], pending[4])


def fold_pending_into_current(device=None):
    """CUDA-graph form of the hand-over: copies every tensor of the pending pyramid (next batch) INTO the tensors of the
    pyramid the current forward used (same shapes: all sizes follow from B, the level sizes and nsample), after joining
    the side stream.  Captured at the end of a step, every replay then reads the pyramid the previous replay built, at
    fixed addresses.  The current pyramid is kept alive for the life of the process (the graph owns its addresses)."""
    key, _ = _side_stream(torch.cuda.current_device() if device is None else device)
    pending = _pending.pop(key, None)
    if pending is None:
        raise RuntimeError("fold_pending_into_current: nothing was prefetched")
    if pending[2] is not None:
        torch.cuda.current_stream().wait_event(pending[2])
    cur, new = list(cache.entries.values()), list(pending[1].values())
    if len(cur) != len(new) or any(type(a) is not type(b) for a, b in zip(cur, new)):
        raise RuntimeError("fold_pending_into_current: the pending pyramid has another structure than the current one")
    a_cur, a_new = getattr(cache, "arena", None), pending[4]
    whole = (a_cur is not None and a_cur.buf is not None and a_new.buf is not None and a_cur.need == a_new.need
             and a_cur.off == a_new.off == a_cur.need)  # both pyramids lie entirely inside equal blocks
    with torch.no_grad():
        for a, b in zip(cur, new):
            for ta, tb in zip(_tensors(a), _tensors(b)):
                if (ta is None) != (tb is None) or (ta is not None and ta.shape != tb.shape):
                    raise RuntimeError("fold_pending_into_current: tensor shapes differ between the two pyramids")
                if ta is not None and not whole:
                    ta.copy_(tb)
        if whole:
            a_cur.buf.copy_(a_new.buf)
    _persistent.append((cur, pending[3], a_cur))


def join(device=None):
    """Makes the current stream wait for everything `prebuild` enqueued (required before a graph capture ends)."""
    if not torch.cuda.is_available():
        return
    key = torch.cuda.current_device() if device is None or torch.device(device).index is None else torch.device(device).index
    final = _final_events.pop(key, None)
    if final is not None:
        torch.cuda.current_stream().wait_event(final)
