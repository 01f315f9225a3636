"""CPU suite, part 4: the N>1 path on world_size-2 gloo — rendezvous, per-rank shards of the patch stream,
DDP gradient averaging configured like the reference, max-over-ranks timing reduction."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from deep3dpointclouddenoising_b200 import distributed, synthetic
    from deep3dpointclouddenoising_b200.models.heads import MultiDimHeadResNet
    from deep3dpointclouddenoising_b200.models.losses import MaskedL1Loss
    from deep3dpointclouddenoising_b200.utils.config import runtime
    runtime.cpu_modules = True  # host-side test of the data-parallel plumbing: conv / BN blocks as stock torch modules
    r, w, _ = distributed.init("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)  # same initial weights on every rank, like DDP's initial broadcast would give
    head = MultiDimHeadResNet(3, 8, 0.025, [4, 4, 4, 4, 4])
    block = torch.nn.Sequential(head.up_conv3, head.head)  # the CPU-runnable tail of the U-Net: conv/BN/ReLU -> 3 dims
    net = distributed.wrap(block)
    assert isinstance(net, torch.nn.parallel.DistributedDataParallel) and not net.broadcast_buffers
    pts, mask, feats, offs = synthetic.make_batch(distributed.shard_seed(rank, 0), 2, 64)
    x = torch.from_numpy(np.tile(feats, (1, 6, 1))[:, :16])  # (B, 2*width, N) stand-in for the concatenated features
    loss = MaskedL1Loss()(net(x).transpose(1, 2), torch.from_numpy(offs), torch.from_numpy(mask))
    loss.backward()
    grads = torch.cat([p.grad.flatten() for p in block.parameters()])
    # the same computation without DDP gives this rank's local gradient
    torch.manual_seed(0)
    head2 = MultiDimHeadResNet(3, 8, 0.025, [4, 4, 4, 4, 4])
    block2 = torch.nn.Sequential(head2.up_conv3, head2.head)
    MaskedL1Loss()(block2(x).transpose(1, 2), torch.from_numpy(offs), torch.from_numpy(mask)).backward()
    local = torch.cat([p.grad.flatten() for p in block2.parameters()])
    # the DDP-free path: grads as views of one flat bucket + a single all-reduce
    torch.manual_seed(0)
    head3 = MultiDimHeadResNet(3, 8, 0.025, [4, 4, 4, 4, 4])
    block3 = torch.nn.Sequential(head3.up_conv3, head3.head)
    bucket = distributed.FlatGradAllReduce(block3.parameters())
    bucket.zero()
    MaskedL1Loss()(block3(x).transpose(1, 2), torch.from_numpy(offs), torch.from_numpy(mask)).backward()
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in block3.parameters())  # still views of the bucket
    bucket.reduce()
    flat_grads = torch.cat([p.grad.flatten() for p in block3.parameters()])
    times = distributed.max_over_ranks([10.0 + rank, 5.0 - rank], "cpu")
    distributed.barrier()
    torch.save({"ddp": grads, "local": local, "flat": flat_grads, "times": times, "seed": distributed.shard_seed(rank, 0),
                "points": torch.from_numpy(pts)}, os.path.join(out_dir, f"rank{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_two_rank_gradient_allreduce_and_sharding(tmp_path):
    world = 2
    for attempt in range(3):  # a rendezvous can lose the race for a just-released port: retry on a fresh one
        try:
            mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
            break
        except Exception:
            if attempt == 2:
                raise
    r0, r1 = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert torch.equal(r0["ddp"], r1["ddp"])  # every rank holds the same averaged gradient after the all-reduce
    torch.testing.assert_close(r0["ddp"], (r0["local"] + r1["local"]) / 2, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(r0["flat"], r0["ddp"], rtol=1e-6, atol=1e-8)  # flat-bucket all-reduce == DDP
    assert torch.equal(r0["flat"], r1["flat"])
    assert not torch.equal(r0["local"], r1["local"])  # the ranks really worked on different shards
    assert r0["seed"] != r1["seed"] and not torch.equal(r0["points"], r1["points"])
    assert r0["times"] == r1["times"] == [11.0, 5.0]  # max over ranks, element-wise


def test_flat_parameters_match_per_tensor_adam_and_clipping():
    """distributed.FlatParameters: Adam + clip_grad_norm_ on the one flat parameter == on the individual tensors."""
    from deep3dpointclouddenoising_b200 import distributed

    def make():
        torch.manual_seed(3)
        return torch.nn.Sequential(torch.nn.Conv1d(3, 10, 1), torch.nn.BatchNorm1d(10), torch.nn.ReLU(),
                                   torch.nn.Conv1d(10, 3, 1))  # sizes 30, 10, 10, 10, 30, 3: exercises the padding

    x = torch.randn(4, 3, 50)
    ref, flat_model = make(), make()
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-2, weight_decay=1e-3)
    flat = distributed.FlatParameters(flat_model)
    assert all(p.data_ptr() % 16 == 0 for p in flat_model.parameters())
    opt_flat = torch.optim.Adam([flat.param], lr=1e-2, weight_decay=1e-3)
    for _ in range(5):
        opt_ref.zero_grad(set_to_none=True)
        ref(x).abs().sum().backward()
        n_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        opt_ref.step()
        flat.zero()
        flat_model(x).abs().sum().backward()
        flat.reduce()  # world size 1: no-op
        n_flat = torch.nn.utils.clip_grad_norm_([flat.param], 0.5)
        opt_flat.step()
        torch.testing.assert_close(n_flat, n_ref, rtol=1e-5, atol=1e-7)
    # Adam divides by sqrt(v): for a parameter whose gradient is ~0 the last-bit differences of two reduction orders
    # are amplified to ~1e-6 after 5 steps of size 1e-2 — hence atol 1e-5 (a wrong update rule is off by ~1e-2)
    for a, b in zip(flat_model.parameters(), ref.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-3, atol=1e-5)
    assert set(flat_model.state_dict()) == set(ref.state_dict())


class _FakeUNet(torch.nn.Module):
    """CPU stand-in with the stage layout of the U-Net (backbone.layer1..4, segmentation_head) and the same markers."""

    def __init__(self):
        super().__init__()
        bb = torch.nn.Module()
        bb.conv1 = torch.nn.Conv1d(3, 6, 1)
        for i in range(1, 5):
            setattr(bb, f"layer{i}", torch.nn.Conv1d(6, 6, 1))
        self.backbone = bb
        self.segmentation_head = torch.nn.Sequential(torch.nn.Conv1d(12, 5, 1), torch.nn.Conv1d(5, 3, 1))

    def forward(self, x):
        from deep3dpointclouddenoising_b200.fused import stage_marker
        f = self.backbone.conv1(x)
        skips = []
        for i in range(1, 5):
            f = getattr(self.backbone, f"layer{i}")(stage_marker(f, f"layer{i}")).relu()
            skips.append(f)
        last = stage_marker(skips[-1], "head")
        return self.segmentation_head(torch.cat([last, skips[0]], 1))


def _overlap_worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from deep3dpointclouddenoising_b200 import distributed, fused
    distributed.init("gloo")
    torch.manual_seed(0)
    model = _FakeUNet()
    flat = distributed.FlatParameters(model).overlap_with_backward(model)
    calls = []
    inner = fused._stage_callback[0]
    fused.set_stage_callback(lambda tag: (calls.append((tag, flat._done_upto)), inner(tag)))
    torch.manual_seed(100 + rank)
    x = torch.randn(4, 3, 32)
    for step in range(2):
        flat.zero()
        model(x).square().sum().backward()
        flat.reduce()
    fused.set_stage_callback(None)
    torch.manual_seed(0)
    ref = _FakeUNet()
    ref(x).square().sum().backward()
    local = torch.cat([torch.nn.functional.pad(p.grad.flatten(), (0, (-p.numel()) % 4)) for p in ref.parameters()])
    torch.save({"flat": flat.flat.clone(), "local": local, "calls": calls}, os.path.join(out_dir, f"ov{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_gradient_allreduce_overlapped_with_backward(tmp_path):
    """distributed.FlatParameters.overlap_with_backward: the bucket goes out in slices as backward leaves the stages
    (head first, then layer4, 3, 2, the rest in reduce()); the result equals the plain average of the local gradients."""
    world = 2
    for attempt in range(3):
        try:
            mp.spawn(_overlap_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
            break
        except Exception:
            if attempt == 2:
                raise
    r0, r1 = [torch.load(tmp_path / f"ov{r}.pt") for r in range(world)]
    assert torch.equal(r0["flat"], r1["flat"])
    torch.testing.assert_close(r0["flat"], (r0["local"] + r1["local"]) / 2, rtol=1e-5, atol=1e-7)
    tags = [t for t, _ in r0["calls"]]
    assert tags[:4] == ["head", "layer4", "layer3", "layer2"], tags  # backward order; layer1's marker has no slice of its own
    ends = [e for _, e in r0["calls"][:4]]
    assert ends == sorted(ends, reverse=True) and ends[0] == r0["flat"].numel()  # contiguous slices from the tail down
