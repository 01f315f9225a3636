"""ctypes binding of libd3d_b200.so (C ABI declared in include/d3d_b200.h).

The library is the product: there is NO CPU or PyTorch fallback.  If the shared object is missing the
import of any op raises; if a tensor is not on a CUDA device the op raises "CPU not supported" like the
reference extension does (u_net_arch/pt_custom_ops/_ext_src/src/group_points.cpp:36).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libd3d_b200.so")

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
_ll = ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/d3d_b200.h one to one (tests/test_abi.py checks the
# header against this table and against the symbols the .so exports).
SIGNATURES = {
    "d3d_abi_version": (_i, []),
    "d3d_error_string": (ctypes.c_char_p, [_i]),
    "d3d_kernel_launches": (ctypes.c_longlong, []),
    "d3d_ball_query_workspace_bytes": (_sz, [_i, _i, _i]),
    "d3d_ball_query": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "d3d_nearest_query_workspace_bytes": (_sz, [_i]),
    "d3d_nearest_query": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "d3d_grid_subsample_workspace_bytes": (_sz, [_i, _i]),
    "d3d_grid_subsample": (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "d3d_group_points": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "d3d_group_points_grad_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "d3d_group_points_grad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "d3d_group_points_grad_atomic": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "d3d_inverse_map_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "d3d_build_inverse_map": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "d3d_cm_to_cl": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "d3d_cl_to_cm": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "d3d_pospool_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "d3d_pospool_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "d3d_spatial_order": (_i, [_vp, _i, _i, _vp, _vp]),
    "d3d_pospool_tile_plan_bytes": (_sz, [_i, _i, _i, _i]),
    "d3d_pospool_tile_plan": (_i, [_vp] * 4 + [_i] * 4 + [_vp, _sz, _vp]),
    "d3d_pospool_tiles_fwd": (_i, [_vp] * 8 + [_i] * 5 + [_f, _i, _vp, _vp]),
    "d3d_pospool_tiles_debug_timing": (None, [_vp]),
    "d3d_pospool_scatter_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "d3d_pospool_scatter_bwd": (_i, [_vp] * 8 + [_i] * 5 + [_f, _i, _vp, _vp, _sz, _vp]),
    "d3d_pseudogrid_fwd": (_i, [_vp] * 8 + [_i] * 6 + [_f, _i, _i, _vp, _vp]),
    "d3d_pseudogrid_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "d3d_pseudogrid_bwd": (_i, [_vp] * 11 + [_i] * 6 + [_f, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "d3d_gather_max_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "d3d_gather_max_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "d3d_nearest_gather_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "d3d_nearest_gather_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "d3d_nn_workspace_bytes": (_sz, [_i]),
    "d3d_nn_sqdist": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "d3d_chamfer_workspace_bytes": (_sz, [_i, _i]),
    "d3d_chamfer_l2": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "d3d_voxel_ids": (_i, [_vp, _i, _f, _f, _f, _f, _i, _i, _i, _vp, _vp]),
    "d3d_voxel_barycentres": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "d3d_radius_patches_workspace_bytes": (_sz, [_i, _i, _i]),
    "d3d_radius_patches": (_i, [_vp, _i, _vp, _i, _f, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "d3d_radius_patches_tier": (_i, [_vp, _i, _vp, _i, _f, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "d3d_vote_mean": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "d3d_bn_act_workspace_bytes": (_sz, [_i]),
    "d3d_bn_act_fwd": (_i, [_vp] * 6 + [_i, _i, _i, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "d3d_bn_act_bwd": (_i, [_vp] * 7 + [_i] * 5 + [_vp] * 5 + [_sz, _vp]),
    "d3d_bn_act_cl_fwd": (_i, [_vp] * 7 + [_ll, _i, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "d3d_linear_small_k": (_i, [_vp, _vp, _ll, _ll, _vp, _ll, _i, _i, _vp, _vp]),
    "d3d_linear_small_n": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp]),
    "d3d_wgrad_small_workspace_bytes": (_sz, [_i]),
    "d3d_wgrad_small": (_i, [_vp, _vp, _ll, _i, _i, _vp, _ll, _ll, _i, _vp, _sz, _vp]),
    "d3d_gemm_row_tiles": (_i, [_ll]),
    "d3d_gemm_tf32": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp, _vp]),
    "d3d_gemm_tf32_act": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "d3d_wgrad_workspace_bytes": (_sz, [_ll, _i, _i]),
    "d3d_wgrad_tf32": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i, _vp, _sz, _vp]),
    "d3d_bn_finalize": (_i, [_vp, _ll, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "d3d_bn_apply_cl": (_i, [_vp] * 6 + [_ll, _i, _i, _vp, _vp]),
    "d3d_bn_act_cl_bwd": (_i, [_vp] * 7 + [_ll, _i, _i, _i] + [_vp] * 4 + [_i, _vp, _sz, _vp]),
}

_lib = None


def load():
    """Loads the shared library once; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m deep3dpointclouddenoising_b200.build_ext` "
                "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class D3DError(RuntimeError):
    pass


def check(code, what):
    if code != 0:
        msg = load().d3d_error_string(int(code)).decode()
        raise D3DError(f"{what} failed with code {code}: {msg}")
