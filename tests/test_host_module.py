"""CPU tests of the host-side mirror of the reference interface (no kernel runs)."""
import pytest
import torch

from oracle import pillars_oracle as po
from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg


def make(C=384, **kw):
    cfg = default_cfg(device="cpu", patch_feature_dim=C, **kw)
    return PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, C]},
                               scatter={"in_channels": C, "output_shape": [28, 28]}, local_rank=0)


def test_state_dict_is_interchangeable_with_reference_names():
    enc = make()
    ref = po.OraclePointPillarsEncoder(po.GridSpec())
    assert list(enc.state_dict().keys()) == list(ref.state_dict().keys())
    sd, _ = po.synth_weights(4)
    assert enc.load_state_dict(sd, strict=True).missing_keys == []
    ref.load_state_dict(enc.state_dict(), strict=True)
    bn = enc.voxel_encoder.pfn_layers[1].norm
    assert isinstance(bn, torch.nn.BatchNorm1d) and bn.eps == 1e-3 and bn.momentum == 0.01
    assert enc.voxel_encoder.pfn_layers[0].linear.bias is None


def test_cfg_mapping_follows_reference_ctor():
    enc = make(max_num_points_per_voxel=128, max_num_voxels=(500, 784))
    assert enc.point_cloud_range == [0, 0, 0, 224, 224, 100] and enc.voxel_size == [8, 8, 100]
    assert enc.max_num_points == 128 and enc.max_voxels == [500, 784]
    enc.train()
    assert enc._grid().max_voxels == 500
    enc.eval()
    assert enc._grid().max_voxels == 784


def test_syncbn_conversion_finds_batchnorm_children():
    enc = torch.nn.SyncBatchNorm.convert_sync_batchnorm(make())
    assert isinstance(enc.voxel_encoder.pfn_layers[0].norm, torch.nn.SyncBatchNorm)


def test_rejects_wrong_inputs_loudly():
    enc = make().eval()
    with pytest.raises(RuntimeError, match="CUDA only"):
        enc(torch.zeros(1, 10, 3))
    with pytest.raises(TypeError):
        enc._pack(torch.zeros(1, 10, 3, dtype=torch.float64))
    with pytest.raises(ValueError):
        enc._pack(torch.zeros(10, 3))
    with pytest.raises(NotImplementedError):
        cfg = default_cfg(device="cpu")
        PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [32, 384]}, scatter={"in_channels": 384, "output_shape": [28, 28]})


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pixelspointspolygons_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.P3PError, match="no CPU or eager fallback"):
        _lib.lib()


def test_product_package_never_imports_the_oracle():
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pixelspointspolygons_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "libp3p_oracle" not in txt, f
