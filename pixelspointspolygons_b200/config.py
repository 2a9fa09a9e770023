"""Minimal stand-in for the reference's Hydra/OmegaConf config object.

The encoder reads `cfg.experiment.encoder.*`, `cfg.host.device`, `cfg.run_type.logging` and
`cfg.experiment.lidar_dropout` (R:pixelspointspolygons/models/pointpillars/pointpillars_o3d.py:34-47,54;
R:.../fusion_layers/early_fusion_vit.py:113).  A real OmegaConf DictConfig works unchanged; `AttrDict` gives the
same attribute + mapping access where hydra/omegaconf are not installed (tests, bench).
"""
from __future__ import annotations


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(obj):
        if isinstance(obj, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in obj.items()})
        return obj


def default_cfg(device="cuda", in_size=224, voxel=(8, 8, 100), max_num_points_per_voxel=64, max_num_voxels=(784, 784),
                patch_feature_dim=384, patch_size=8, lidar_dropout=None, logging="WARNING", **encoder_extra):
    """Values of R:config/encoder/{pointpillars_vit,early_fusion_vit}.yaml."""
    feat = in_size // patch_size
    enc = dict(
        in_size=in_size, in_height=in_size, in_width=in_size,
        in_voxel_size=dict(x=voxel[0], y=voxel[1], z=voxel[2]),
        max_num_points_per_voxel=max_num_points_per_voxel,
        max_num_voxels=dict(train=max_num_voxels[0], test=max_num_voxels[1]),
        patch_size=patch_size, patch_feature_size=feat, patch_feature_height=feat, patch_feature_width=feat,
        patch_feature_dim=patch_feature_dim, num_patches=feat * feat,
    )
    enc.update(encoder_extra)
    return AttrDict.wrap(dict(
        host=dict(device=device),
        run_type=dict(logging=logging),
        experiment=dict(encoder=enc, lidar_dropout=lidar_dropout),
    ))


def _resolve(root, path, value):
    """OmegaConf-style relative interpolations `${.a}`, `${..a.b}` inside strings (what the reference's encoder yamls use)."""
    import re

    def lookup(m):
        expr = m.group(1)
        dots = len(expr) - len(expr.lstrip("."))
        if dots == 0:
            node, keys = root, expr.split(".")
        else:
            node, keys = root, path[:len(path) - dots]
            for k in keys:
                node = node[k]
            keys = expr.lstrip(".").split(".")
        for k in keys:
            node = node[k]
        return node

    if not isinstance(value, str) or "${" not in value:
        return value
    whole = re.fullmatch(r"\$\{([^${}]+)\}", value)
    if whole:
        return _resolve(root, path, lookup(whole))
    return re.sub(r"\$\{([^${}]+)\}", lambda m: str(_resolve(root, path, lookup(m))), value)


def _resolve_tree(root, node=None, path=()):
    node = root if node is None else node
    for k, v in list(node.items()):
        if isinstance(v, dict):
            _resolve_tree(root, v, path + (k,))
        else:
            node[k] = _resolve(root, path + (k,), v)
    return root


def cfg_from_encoder_yaml(path, device="cuda", out_path=".", decoder_in_feature_dim=256, lidar_dropout=None, logging="WARNING",
                          **encoder_overrides):
    """The cfg object the reference composes with Hydra for one of its encoder yamls (R:config/encoder/*.yaml), built with
    PyYAML alone: `cfg.experiment.encoder` = the file, with its relative `${...}` interpolations resolved against the
    surrounding nodes the encoders read (`experiment.model.decoder.in_feature_dim`, `experiment.dataset.out_path`,
    `host.device`, `run_type.logging`, `experiment.lidar_dropout`)."""
    import yaml

    with open(path) as f:
        enc = yaml.safe_load(f)
    enc.update(encoder_overrides)
    tree = dict(host=dict(device=device), run_type=dict(logging=logging),
                experiment=dict(encoder=enc, model=dict(decoder=dict(in_feature_dim=decoder_in_feature_dim)),
                                dataset=dict(out_path=out_path), lidar_dropout=lidar_dropout))
    return AttrDict.wrap(_resolve_tree(tree))
