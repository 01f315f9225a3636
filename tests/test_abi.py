"""CPU suite, part 2: the C-ABI library builds, loads and exports every symbol include/d3d_b200.h declares
(no compute calls — there is no GPU here), and the host-side checks fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from deep3dpointclouddenoising_b200 import build_ext, _lib
    build_ext.build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "d3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d3d_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from deep3dpointclouddenoising_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in d3d_b200.h but not exported by libd3d_b200.so"
    assert sorted(_lib.SIGNATURES) == declared


def test_abi_version_and_error_strings(lib):
    assert lib.d3d_abi_version() == 4
    assert lib.d3d_error_string(0) == b"ok"
    assert b"workspace" in lib.d3d_error_string(-3)


def test_workspace_queries_are_pure_host_functions(lib):
    assert lib.d3d_ball_query_workspace_bytes(16, 8192, 8192) >= 16 * 4 + 16 * 8192 * (16 + 8)
    assert lib.d3d_grid_subsample_workspace_bytes(16, 8192) == 0  # sorted in shared memory
    assert lib.d3d_grid_subsample_workspace_bytes(2, 20000) == 2 * 32768 * 8
    assert lib.d3d_inverse_map_workspace_bytes(16, 8192, 8192, 52) >= 16 * 8192 * 52 * 4
    assert lib.d3d_pseudogrid_bwd_workspace_bytes(16, 8192, 72, 15) > 0


def test_bad_arguments_are_rejected_without_touching_the_device(lib):
    assert lib.d3d_ball_query(None, None, None, None, 1, 1, 1, 0.1, 4, None, None, None, None, None, 0, None) == -1
    assert lib.d3d_group_points(None, None, 1, 1, 1, 1, 1, None, None) == -1


def test_cpu_tensors_raise_like_the_reference_extension():
    from deep3dpointclouddenoising_b200.pt_custom_ops import _ext
    f = torch.zeros(1, 3, 8)
    idx = torch.zeros(1, 4, 2, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="CPU not supported"):  # group_points.cpp:36
        _ext.group_points(f, idx)
    xyz, m = torch.zeros(1, 8, 3), torch.ones(1, 8, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        _ext.masked_ordered_ball_query(xyz, xyz, m, m, 0.1, 4)


def test_reference_extension_surface():
    from deep3dpointclouddenoising_b200.pt_custom_ops import _ext, pt_utils
    for name in ("group_points", "group_points_grad", "masked_ordered_ball_query", "masked_nearest_query",
                 "masked_grid_subsampling"):  # bindings.cpp:8-14
        assert callable(getattr(_ext, name))
    for name in ("grouping_operation", "masked_ordered_ball_query", "masked_nearest_query", "masked_grid_subsampling",
                 "MaskedQueryAndGroup", "MaskedNearestQueryAndGroup", "MaskedMaxPool", "MaskedUpsample"):
        assert hasattr(pt_utils, name)
