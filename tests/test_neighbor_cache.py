"""CPU suite, part 7: the per-forward neighbour cache (neighbors.py) — keyed on tensor identity AND
version, so an in-place update of the coordinates can never reuse a stale list; the side-stream prebuild is a no-op for
CPU tensors (the kernels themselves refuse them)."""
import torch

from deep3dpointclouddenoising_b200 import neighbors


def test_cache_keys_on_identity_and_version():
    cache = neighbors._Cache()
    xyz, mask = torch.randn(2, 8, 3), torch.ones(2, 8, dtype=torch.int32)
    built = []

    def build():
        built.append(1)
        return object()

    a = cache.get("ball", (xyz, xyz, mask, mask), (0.1, 4), build)
    assert cache.get("ball", (xyz, xyz, mask, mask), (0.1, 4), build) is a and len(built) == 1  # hit
    assert cache.get("ball", (xyz, xyz, mask, mask), (0.2, 4), build) is not a                   # other radius
    assert cache.get("ball", (xyz.clone(), xyz, mask, mask), (0.1, 4), build) is not a           # other tensor
    xyz.add_(1.0)                                                                                 # same storage, new version
    assert cache.get("ball", (xyz, xyz, mask, mask), (0.1, 4), build) is not a
    assert (cache.hits, cache.misses) == (1, 4)
    cache.clear()
    assert not cache.entries
    cache.enabled = False
    n = len(built)
    cache.get("ball", (xyz, xyz, mask, mask), (0.1, 4), build)
    cache.get("ball", (xyz, xyz, mask, mask), (0.1, 4), build)
    assert len(built) == n + 2 and not cache.entries  # disabled: always rebuilt, nothing stored


def test_prebuild_and_join_are_noops_without_cuda_tensors():
    xyz, mask = torch.randn(1, 16, 3), torch.ones(1, 16, dtype=torch.int32)
    neighbors.cache.clear()
    neighbors.prebuild(xyz, mask, 0.1, 4, [(0.05, 4, 0.1, 4, 0.2, 4)], with_csr=True)
    assert not neighbors.cache.entries
    neighbors.join()
