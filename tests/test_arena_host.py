"""CPU suite: ops.Arena — the block the outputs of a neighbourhood-pyramid build are carved from (one-step pipeline,
neighbors.prefetch / fold_pending_into_current) — and the per-item tensor enumeration the hand-over relies on."""
import torch

from deep3dpointclouddenoising_b200 import neighbors, ops


def test_arena_measures_then_carves_aligned_views():
    shapes = [((2, 5, 3), torch.float32), ((2, 5), torch.int32), ((7,), torch.int32), ((0,), torch.int32), ((3, 3), torch.uint8)]
    probe = ops.Arena()
    with probe:
        first = [ops._out(s, d, "cpu") for s, d in shapes]
    assert ops.Arena._current is None
    assert probe.buf is None and probe.need == sum((max(t.numel() * t.element_size(), 0) + 255) // 256 * 256 for t in first)
    arena = ops.Arena(probe.need, "cpu")
    with arena:
        second = [ops._out(s, d, "cpu") for s, d in shapes]
    assert arena.off <= arena.need == probe.need
    base = arena.buf.data_ptr()
    for t, (s, d) in zip(second, shapes):
        assert tuple(t.shape) == s and t.dtype == d
        if t.numel():
            assert base <= t.data_ptr() < base + arena.buf.numel() and (t.data_ptr() - base) % 256 == 0
    # the views alias the block: one copy of the block moves every tensor
    other = ops.Arena(probe.need, "cpu")
    with other:
        third = [ops._out(s, d, "cpu") for s, d in shapes]
    for t in third:
        t.fill_(7)
    arena.buf.copy_(other.buf)
    assert all(bool((t == 7).all()) for t in second if t.numel())
    # outside an arena the helper is a plain allocation
    assert ops._out((4,), torch.float32, "cpu").shape == (4,)


def test_arena_falls_back_when_exhausted():
    arena = ops.Arena(256, "cpu")
    with arena:
        a = ops._out((32,), torch.float32, "cpu")   # 128 bytes -> one 256-byte slot
        b = ops._out((32,), torch.float32, "cpu")   # does not fit any more: a fresh allocation, still counted
    assert a.data_ptr() == arena.buf.data_ptr() and not (arena.buf.data_ptr() <= b.data_ptr() < arena.buf.data_ptr() + 256)
    assert arena.need == 512


def test_hand_over_enumerates_every_tensor_of_an_item():
    idx = torch.zeros(1, 4, 2, dtype=torch.int32)
    nbr = neighbors.NeighborList(idx, idx.clone(), torch.zeros(1, 4, dtype=torch.int32), 4, by_support=idx.clone())
    nbr._csr = (torch.zeros(5, dtype=torch.int32), torch.zeros(8, dtype=torch.int32))
    nbr._plan = torch.zeros(16, dtype=torch.uint8)
    ts = neighbors._tensors(nbr)
    assert len(ts) == 7 and all(t is not None for t in ts)
    sub = neighbors._Subsampled(torch.zeros(1, 2, 3), torch.ones(1, 2, dtype=torch.int32), ())
    assert len(neighbors._tensors(sub)) == 2
    assert neighbors._tensors(neighbors._Order(torch.zeros(1, 4, dtype=torch.int32), ())) [0].shape == (1, 4)


def test_runtime_switch_helpers():
    from deep3dpointclouddenoising_b200.utils.config import runtime, set_deterministic
    old = (runtime.staged_tiles_backward, runtime.deterministic_scatter)
    try:
        set_deterministic(True)
        assert runtime.staged_tiles_backward == 'ordered' and runtime.deterministic_scatter is True
        set_deterministic(False)
        assert runtime.staged_tiles_backward == 'scatter' and runtime.deterministic_scatter is False
    finally:
        runtime.staged_tiles_backward, runtime.deterministic_scatter = old
    # the skinny-layer kernels only take CUDA fp32 tensors: anything else goes to the library path
    x, w = torch.zeros(8, 3), torch.zeros(72, 3)
    assert ops.small_linear_kind(3, 72, x, w) is None
