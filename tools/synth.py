"""Seeded synthetic inputs of SURVEY 8(d): tiles and eval-mode weights shared by the tests, smoke(), bench.py and the
oracle.  Plain numpy / torch generators: neither the product package nor the oracle is imported here."""
from __future__ import annotations

import numpy as np
import torch


def synth_tile(n_points: int, seed: int, size: float = 224.0, zmax: float = 100.0, clustered: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if n_points == 0:
        return np.zeros((0, 3), np.float32)
    xy = rng.uniform(0.0, size, (n_points, 2))
    if clustered:
        k = int(0.3 * n_points)
        centres = rng.uniform(16.0, size - 16.0, (20, 2))
        which = rng.integers(0, 20, k)
        xy[:k] = centres[which] + rng.uniform(-8.0, 8.0, (k, 2))
    z = rng.uniform(0.0, zmax, (n_points, 1))
    pts = np.concatenate([xy, z], 1).astype(np.float32)
    pts[:, :2] = np.minimum(pts[:, :2], np.nextafter(np.float32(size), np.float32(0)))
    pts[:, 2] = np.minimum(pts[:, 2], np.nextafter(np.float32(zmax), np.float32(0)))
    if n_points >= 2:  # MinMaxScaler(0, zmax) guarantees one point at each end (p3_coco.py:87-88)
        pts[rng.integers(0, n_points), 2] = 0.0
        pts[rng.integers(0, n_points), 2] = zmax
    return pts[rng.permutation(n_points)]


def synth_weights(seed: int, feat_channels=(64, 384), in_channels=3, patch=8, img_chans=3):
    """Seeded eval-mode parameters with the reference's state_dict keys."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    chans = [in_channels + 5] + list(feat_channels)
    for i in range(len(chans) - 1):
        units = chans[i + 1] if i == len(chans) - 2 else chans[i + 1] // 2
        fan_in = chans[i]
        p = f"voxel_encoder.pfn_layers.{i}."
        sd[p + "linear.weight"] = torch.randn(units, fan_in, generator=g) * (0.1 if i == 0 else 0.15)
        sd[p + "norm.weight"] = torch.rand(units, generator=g) + 0.5
        sd[p + "norm.bias"] = torch.randn(units, generator=g) * 0.1
        sd[p + "norm.running_mean"] = torch.randn(units, generator=g) * 0.1
        sd[p + "norm.running_var"] = torch.rand(units, generator=g) + 0.5
        sd[p + "norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        # make a few BN scales negative: exercises max-before-affine orderings
        sd[p + "norm.weight"][::7] *= -1.0
    C = chans[-1]
    k = img_chans * patch * patch
    sd_img = {"proj.weight": torch.randn(C, img_chans, patch, patch, generator=g) / (k ** 0.5),
              "proj.bias": torch.randn(C, generator=g) * 0.1}
    return sd, sd_img
