"""Global experiment config with the reference's keys, defaults and strict yaml overlay.

Mirrors u_net_arch/utils/config.py: a module-level attribute dictionary `config` (:4-142) and
`update_config(path)` (:145-156), which copies a yaml file over the defaults and raises ValueError for
any key the defaults do not know.  `easydict` (the reference's container) is not a dependency here;
AttrDict below gives the same attribute + item access.  Reference cfgs/*.yaml files load unchanged.
"""
import copy

import yaml


class AttrDict(dict):
    """dict with attribute access, nested dicts converted on assignment (what easydict provides)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


_DEFAULTS = {
    # experiment options (config.py:9-24)
    "experiment_name": "", "noise_level": -1, "outlier_percentage": -1, "epoch_model_used": -1, "loss": "l1",
    "jitter": 0, "norm": 0, "GAN": 0, "load_path_generator": "", "load_path_discriminator": "",
    "head_discriminator": "None", "freeze_gen": 0, "architecture": "U-Net", "noise_type": "gaussian",
    "sample_Dl_patches": 0.05, "fourier_features": 0,
    # training options (:30-41)
    "epochs": 50, "start_epoch": 1, "base_learning_rate": 0.01, "lr_scheduler": "step", "optimizer": "sgd",
    "warmup_epoch": 5, "warmup_multiplier": 100, "lr_decay_steps": 20, "lr_decay_rate": 0.7, "weight_decay": 0,
    "momentum": 0.9, "grid_clip_norm": -1,
    # model (:45-55)
    "backbone": "resnet", "head": "resnet_cls", "radius": 0.05, "sampleDl": 0.02, "density_parameter": 5.0,
    "nsamples": [], "npoints": [], "width": 144, "depth": 2, "bottleneck_ratio": 2, "bn_momentum": 0.1,
    # data (:60-85)
    "datasets": "modelnet40", "data_root": "", "num_classes": 40, "num_parts": 0, "features": [],
    "input_features_dim": 1, "katz_params": [], "katz_type": "std", "batch_size": 32, "num_points": 5000,
    "num_workers": 4, "x_angle_range": 0.0, "y_angle_range": 0.0, "z_angle_range": 0.0, "scale_low": 2.0 / 3.0,
    "scale_high": 3.0 / 2.0, "noise_std": 0.01, "noise_clip": 0.05, "translate_range": 0.2, "color_drop": 0.2,
    "augment_symmetries": [0, 0, 0],
    # scene segmentation related (:92-93)
    "in_radius": 2.0, "num_steps": 500,
    # io and misc (:98-105)
    "load_path": "", "print_freq": 10, "save_freq": 10, "val_freq": 10, "log_dir": "log", "local_rank": 0,
    "amp_opt_level": "", "rng_seed": 0,
    # local aggregation (:110-141)
    "local_aggregation_type": "pospool",
    "pospool": {"position_embedding": "xyz", "reduction": "sum", "output_conv": False},
    "adaptive_weight": {"weight_type": "dp", "num_mlps": 1, "shared_channels": 1, "weight_softmax": False,
                        "reduction": "avg", "output_conv": False},
    "pointwisemlp": {"feature_type": "dp_df", "num_mlps": 1, "reduction": "max"},
    "pseudo_grid": {"fixed_kernel_points": "center", "KP_influence": "linear", "KP_extent": 1.0,
                    "num_kernel_points": 15, "convolution_mode": "sum", "output_conv": False},
    "attention": {"type": "Non-local"},
}

config = AttrDict(copy.deepcopy(_DEFAULTS))

# Not in the reference: selects the arithmetic of the PseudoGrid contraction.
#   'fp32' CUDA cores (parity tolerance 1e-5) | 'bf16' tcgen05 tensor cores (separate tolerance)
# Kept outside the yaml-checked key set so that reference yaml files stay valid and unknown keys still raise.
#   fused_batchnorm: BatchNorm1d (+ReLU, +residual) through csrc/batchnorm.cu instead of cuDNN/ATen (row f4)
#   channel_last: fused ops return channel-last views and the 1x1 convolutions run as row-major GEMMs, so the
#     network never transposes between the convolutions and the aggregations (fused.py, models/blocks.py)
#   prefetch_neighbors: the backbone enqueues all neighbourhood structures of a forward on a side stream (neighbors.py)
#   grads_in_place: backward kernels ADD parameter gradients straight into existing .grad buffers and return None to
#     autograd (no per-parameter accumulation kernels).  Only valid when nothing hooks the gradients (no DDP reducer):
#     distributed.FlatParameters-style training loops switch it on, the default is off.
#   staged_tiles: PosPool forward runs as the staged-tile tensor-core kernel (csrc/pospool_tiles.cu: 128 spatially
#     adjacent rows per CTA, cp.async.bulk staging of the neighbour-row union, tcgen05 contraction) where it is the
#     faster one: self queries (M == N; measured on B200: 189 vs 268 us at the first level, tools/time_pospool.py).
#     'always' forces it for every shape, False restores the per-query gather kernel everywhere.
#   own_gemm: the 1x1 convolutions (forward and data gradient) run on the package's TMA + tcgen05 TF32 GEMM
#     (csrc/gemm.cu), whose epilogue also emits the BatchNorm statistics; False: cuBLAS through torch.
#   own_wgrad: weight gradients of the 1x1 convolutions on the package's tensor-core GEMM (d3d_wgrad_tf32) for layers with
#     at least own_wgrad_min_rows rows; False (default): torch's batched split-K bmm (cuBLAS).  Measured on B200: the own
#     kernel is at parity on the memory-bound first levels (22 / 37 / 42 us against 20 / 30 / 40 us) and 2-3x slower on
#     the deep levels (128 x 160 tiles re-read the operands through L2; cuBLAS uses 256-wide 2-SM tiles), and the step
#     is 0.1-0.4 ms slower with it — the weight gradients run on a side stream either way.
#   wgrad_side_stream: with grads_in_place, the weight-gradient GEMMs run on a side stream (models/blocks.py); the
#     training loop joins them with distributed.FlatParameters.reduce() / blocks.join_weight_grads() before the optimiser.
#   staged_tiles_backward: 'scatter' (default): PosPool backward as the transposed forward tile on the tensor cores, partial
#     sums added with float atomics like the reference's backward (164 us against 322 us at the first level, step 7.9 ms);
#     'ordered': the same tiles store their partial rows and a second kernel adds them per support row in ascending tile
#     order — no float atomics, bit-reproducible (247 us, step 8.2 ms); False: the atomic-free segmented reduction over the
#     inverse map (322 us, step 8.6 ms).  A support-tile tensor-core form without atomics was measured at 838 us and removed.
#   fold_eval_batchnorm: inference (eval mode, no grad): conv + BatchNorm (+ residual) (+ ReLU) as ONE GEMM — the BatchNorm
#     scale is folded into the weights, shift / residual / ReLU ride on the GEMM's store (d3d_gemm_tf32_act).
#   own_small_linear: the 3-channel input / output convolutions on the fp32 streaming kernels of csrc/linear_small.cu
#     (forward, data and weight gradients) instead of cuBLAS.
#   deterministic_scatter: False — _ext.group_points_grad adds with shared-memory float atomics like the reference's
#     kernel (1.04 ms at B=16 x 8192, C=72, ns=52; the reference's kernel: 23.7 ms); True: fixed-order sums over an inverse
#     map (7.8 ms, bit-reproducible).
#   cpu_modules: False — the 1x1-convolution / BatchNorm blocks raise on CPU tensors (no CPU path in the product); the CPU
#     oracle (oracle/cpu_model.py) sets it while it drives the module tree with the reference's formulas.
runtime = AttrDict({"pseudo_grid_precision": "fp32", "fused_batchnorm": True, "channel_last": True,
                    "prefetch_neighbors": True, "grads_in_place": False, "staged_tiles": True,
                    "staged_tiles_backward": "scatter", "own_gemm": True, "wgrad_side_stream": True,
                    "own_wgrad": False, "own_wgrad_min_rows": 0, "cpu_modules": False,
                    "deterministic_scatter": False, "own_small_linear": True,
                    "fold_eval_batchnorm": True})


def set_deterministic(on=True):
    """Bit-reproducible training steps: every float reduction in a fixed order — PosPool backward as tensor-core tiles with
    the ordered second pass, _ext.group_points_grad over the inverse map.  The defaults use float atomics in those two
    places (like the reference's own backward) and are ~4 % faster."""
    runtime.staged_tiles_backward = 'ordered' if on else 'scatter'
    runtime.deterministic_scatter = bool(on)


def reset_config():
    """Back to the defaults (the reference has no such call; tests need it because `config` is global)."""
    config.clear()
    for k, v in copy.deepcopy(_DEFAULTS).items():
        config[k] = v
    return config


def update_config(config_file):
    with open(config_file) as f:
        overlay = yaml.load(f, Loader=yaml.FullLoader)
    for key, value in (overlay or {}).items():
        if key not in config:
            raise ValueError("{} key must exist in config.py".format(key))
        if isinstance(value, dict):
            for sub_key, sub_value in value.items():
                config[key][sub_key] = sub_value
        else:
            config[key] = value


def apply_train_geometry(cfg, in_radius=0.05):
    """The geometry train_dist.py imposes on every yaml before building the model (train_dist.py:125-137):
    in_radius 0.05, sampleDl = in_radius/32, radius = max(in_radius*sqrt(3)/32, 0.025),
    nsamples [52,39,32,26,26], npoints [N/4, N/16, N/32, N/128] from the yaml's num_points."""
    import numpy as np
    cfg.in_radius = in_radius
    cfg.sampleDl = cfg.in_radius / 32
    cfg.radius = max(cfg.in_radius * np.sqrt(3) / 32, 0.025)
    cfg.nsamples = [52, 39, 32, 26, 26]
    cfg.npoints = [max(int(cfg.num_points / d), 1) for d in (4.0, 16.0, 32.0, 128.0)]
    return cfg
