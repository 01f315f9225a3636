"""Cross-step pipelining of the neighbourhood pyramid (neighbors.prefetch / adopt / fold_pending_into_current): the
training arithmetic must not change — same losses and gradients as building the pyramid inside forward."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _bit_reproducible_backward():
    """These tests compare training steps bit for bit (to 1e-6): PosPool backward without float atomics."""
    from deep3dpointclouddenoising_b200.utils.config import runtime
    old, runtime.staged_tiles_backward = runtime.staged_tiles_backward, False
    yield
    runtime.staged_tiles_backward = old


def _setup(seed=0):
    import bench
    from deep3dpointclouddenoising_b200 import synthetic
    torch.manual_seed(seed)
    model, criterion, _ = bench.build_model("pospool", 2048)
    dev = torch.device("cuda:0")
    model = model.to(dev)
    batches = [[torch.from_numpy(a).to(dev) for a in synthetic.make_batch(100 + s, 2, 2048)] for s in range(3)]
    return model, criterion, batches


def _step(model, criterion, batch, nxt=None):
    model.zero_grad(set_to_none=True)
    pts, mask, feats, offs = batch
    loss = criterion(model(pts, mask, feats).transpose(1, 2), offs, mask)
    if nxt is not None:
        model.prefetch_neighbors(nxt[0], nxt[1])
    loss.backward()
    g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    return float(loss), g.clone()


def test_prefetched_pyramid_is_adopted_and_equal(monkeypatch):
    from deep3dpointclouddenoising_b200 import neighbors
    model, criterion, batches = _setup()
    plain = [_step(model, criterion, b) for b in batches]
    adopted = []
    real_adopt = neighbors.adopt
    monkeypatch.setattr(neighbors, "adopt", lambda xyz, mask: adopted.append(real_adopt(xyz, mask)) or adopted[-1])
    piped = []
    for k, b in enumerate(batches):
        piped.append(_step(model, criterion, b, batches[k + 1] if k + 1 < len(batches) else None))
    torch.cuda.synchronize()
    assert adopted == [False, True, True]  # the first step builds its own pyramid, the next two find theirs prefetched
    for (l0, g0), (l1, g1) in zip(plain, piped):  # BatchNorm running statistics do not enter training-mode outputs
        assert abs(l0 - l1) <= 1e-6 * abs(l0)
        torch.testing.assert_close(g1, g0, rtol=1e-4, atol=1e-6 * float(g0.abs().max()))


def test_mismatched_tensors_are_not_adopted():
    from deep3dpointclouddenoising_b200 import neighbors
    model, criterion, batches = _setup()
    model.prefetch_neighbors(batches[1][0], batches[1][1])
    l_other, _ = _step(model, criterion, batches[0])  # another batch: the pending pyramid must be dropped, not used
    l_ref, _ = _step(model, criterion, batches[0])
    assert abs(l_other - l_ref) <= 1e-6 * abs(l_ref)
    assert not neighbors._pending


def test_graph_replay_with_folded_pyramid_matches_eager():
    """The captured form: the step reads a pyramid at fixed addresses and overwrites it with the next batch's at its end."""
    from deep3dpointclouddenoising_b200 import neighbors
    model, criterion, batches = _setup()
    model.train()
    eager = []
    with torch.no_grad():
        state = {k: v.clone() for k, v in model.state_dict().items()}
    for b in batches:  # no optimiser: the weights stay, only the running statistics move (not used in training mode)
        eager.append(_step(model, criterion, b)[0])
    model.load_state_dict(state)
    cur = [t.clone() for t in batches[0]]
    nxt = [t.clone() for t in batches[1]]
    params = [p for p in model.parameters()]
    for p in params:
        p.grad = torch.zeros_like(p)

    def step():
        for p in params:
            p.grad.zero_()
        loss = criterion(model(cur[0], cur[1], cur[2]).transpose(1, 2), cur[3], cur[1])
        model.prefetch_neighbors(nxt[0], nxt[1])
        loss.backward()
        neighbors.fold_pending_into_current()
        for d, s in zip(cur, nxt):
            d.copy_(s)
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up off the default stream, as capture requires
        model.prefetch_neighbors(cur[0], cur[1])
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for d, s in zip(cur, batches[0]):
        d.copy_(s)
    for d, s in zip(nxt, batches[1]):
        d.copy_(s)
    model.prefetch_neighbors(cur[0], cur[1])
    neighbors.settle()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = step()
    got = []
    for k in range(3):
        for d, s in zip(nxt, batches[(k + 1) % 3]):
            d.copy_(s)
        graph.replay()
        got.append(float(loss))
    np.testing.assert_allclose(got, eager, rtol=1e-5)
