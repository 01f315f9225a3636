"""CPU suite, part 3: config overlay semantics, model construction and state-dict names (host logic)."""
import glob
import os

import pytest

from deep3dpointclouddenoising_b200.utils import config as cfgmod

PKG_CFGS = os.path.join(os.path.dirname(cfgmod.__file__), "..", "cfgs")
REF_CFGS = "/root/reference/u_net_arch/cfgs"


@pytest.fixture(autouse=True)
def fresh_config():
    cfgmod.reset_config()
    yield
    cfgmod.reset_config()


def test_defaults_match_reference_config():
    c = cfgmod.config
    assert c.local_aggregation_type == 'pospool' and c.pospool.position_embedding == 'xyz' and c.pospool.reduction == 'sum'
    assert c.pseudo_grid.num_kernel_points == 15 and c.pseudo_grid.KP_influence == 'linear'
    assert c.density_parameter == 5.0 and c.width == 144 and c.bn_momentum == 0.1


def test_unknown_key_raises_like_reference(tmp_path):
    p = tmp_path / "bad.yaml"
    p.write_text("not_a_key: 1\n")
    with pytest.raises(ValueError, match="key must exist in config.py"):
        cfgmod.update_config(str(p))


def test_shipped_configs_load_and_build():
    from deep3dpointclouddenoising_b200.models import build_offset_regression
    expected = {"l1.yaml": 18434307, "l1_pospool.yaml": 18367347}
    for name, n_params in expected.items():
        cfgmod.reset_config()
        cfgmod.update_config(os.path.join(PKG_CFGS, name))
        c = cfgmod.config
        c.num_points = 8192
        cfgmod.apply_train_geometry(c)
        c.input_features_dim = 0
        assert c.npoints == [2048, 512, 256, 64] and c.nsamples == [52, 39, 32, 26, 26] and c.radius == 0.025
        model, criterion = build_offset_regression(c)
        assert sum(p.numel() for p in model.parameters()) == n_params
        keys = model.state_dict().keys()
        assert "backbone.layer1.strided_bottleneck.local_aggregation.local_aggregation_operator.out_transform.0.weight" in keys
        assert "backbone.layer4.bottlneck0.conv1.0.weight" in keys and "segmentation_head.head.3.bias" in keys
        if name == "l1.yaml":
            assert "backbone.la1.local_aggregation_operator.kernel_weights" in keys
            assert "backbone.la1.local_aggregation_operator.K_points" in keys


@pytest.mark.skipif(not os.path.isdir(REF_CFGS), reason="reference tree not present")
def test_reference_yaml_files_load_unchanged():
    known_broken = {"pseudogrid.yaml", "offset_try2.yaml"}  # already invalid in the reference
    loaded = 0
    for path in sorted(glob.glob(os.path.join(REF_CFGS, "*.yaml"))):
        cfgmod.reset_config()
        if os.path.basename(path) in known_broken:
            with pytest.raises(Exception):
                cfgmod.update_config(path)
            continue
        cfgmod.update_config(path)
        loaded += 1
    assert loaded >= 40


def test_kernel_points_match_reference_fixtures():
    import numpy as np
    from deep3dpointclouddenoising_b200.models.utlis import create_kernel_points
    k = create_kernel_points(0.015, 15, 1, 3, 'center')
    assert k.shape == (1, 15, 3)
    ref = "/root/reference/u_net_arch/kernels/dispositions/sk_pt_0.015000_015_center.npy"
    if os.path.exists(ref):
        assert np.array_equal(k, np.load(ref))
    other = create_kernel_points(0.02, 15, 1, 3, 'center')  # not shipped: generated, centre point near the origin
    assert other.shape == (1, 15, 3) and np.linalg.norm(other[0, 0]) < 0.02 * 0.1
    assert np.abs(np.linalg.norm(other[0, 1:], axis=1)).max() < 0.02 * 1.1
