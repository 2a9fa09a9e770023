"""oracle/pillars_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + PyTorch-CPU, fp32) of the reference's LiDAR pillar
encoder and early-fusion front end, i.e. everything
`PointPillarsEncoder.forward` (R:pixelspointspolygons/models/pointpillars/pointpillars_o3d.py:85-107)
and `EarlyFusionViT.forward` up to the concat
(R:pixelspointspolygons/models/fusion_layers/early_fusion_vit.py:96-121) compute.

PARITY UNPINNED.  The reference class is a thin subclass of
`open3d.ml.torch.models.PointPillars` (pointpillars_o3d.py:5,11); the
arithmetic is in the third-party wheel open3d==0.19.0 (R:pyproject.toml:23,
bundling Open3D-ML as open3d._ml3d), which is not under /root/reference and
cannot be installed offline.  The reference has no tests / golden vectors for
this path.  What follows restates the published
algorithm of upstream Open3D-ML `ml3d/torch/models/point_pillars.py`
(PointPillarsVoxelization, PillarFeatureNet, PFNLayer, PointPillarsScatter,
PointPillars.voxelize) and Open3D `cpp/open3d/ml/impl/misc/Voxelize.h`,
following SURVEY.md Appendix A, anchored on the reference's own call sites:

  pointpillars_o3d.py:39-60   cfg -> point_cloud_range / voxel_size / max_num_points / max_voxels
  pointpillars_o3d.py:92-95   voxelize -> voxel_encoder -> middle_encoder(batch_size = x_lidar.shape[0])
  pointpillars_o3d.py:104-107 optional flatten(2).transpose(1,2)
  pointpillars_vit.py:55-64   voxel_encoder = {in_channels: 3, feat_channels: [64, C]}, scatter = {in_channels: C, output_shape}
  early_fusion_vit.py:99-121  image_embed (timm PatchEmbed conv, flatten=False) / lidar dropout / cat(dim=1)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# --------------------------------------------------------------------------
# configuration mirror of pointpillars_o3d.py:39-47 (+ scatter / feat dims)
# --------------------------------------------------------------------------
@dataclass
class GridSpec:
    in_width: float = 224.0
    in_height: float = 224.0
    voxel_size: Tuple[float, float, float] = (8.0, 8.0, 100.0)
    max_num_points: int = 64
    max_voxels: Tuple[int, int] = (784, 784)  # (train, test)
    output_shape: Tuple[int, int] = (28, 28)  # [patch_feature_width, patch_feature_height] -> (ny, nx)
    feat_channels: Tuple[int, ...] = (64, 384)
    in_channels: int = 3
    drop_overflow: bool = False  # see voxelize_ref.c header "Hashes >= batch_hash" (ledger U1)

    @property
    def range_min(self):
        return (0.0, 0.0, 0.0)

    @property
    def range_max(self):
        # point_cloud_range = [0,0,0, in_width, in_height, in_voxel_size.z]  (pointpillars_o3d.py:39-40)
        return (float(self.in_width), float(self.in_height), float(self.voxel_size[2]))


def load_lib():
    """Build (if needed) and load the plain-C voxelizer restatement."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "libp3p_oracle.so")
    src = os.path.join(_HERE, "voxelize_ref.c")
    if not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libp3p_oracle.so"])
    lib = ctypes.CDLL(so)
    i64, p = ctypes.c_int64, ctypes.c_void_p
    lib.p3p_oracle_voxelize.argtypes = [p, i64, i64, p, p, p, i64, i64, i64, p, p, p, p, p]
    lib.p3p_oracle_voxelize.restype = ctypes.c_int
    lib.p3p_oracle_pillarize.argtypes = [p, i64, i64, p, p, p, i64, i64, i64, p, p, p, p]
    lib.p3p_oracle_pillarize.restype = ctypes.c_int64
    _LIB = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------
# A.1  open3d.ml.torch.ops.voxelize (CPU), batch_size == 1
# --------------------------------------------------------------------------
def voxelize_c(points: np.ndarray, grid: GridSpec, max_voxels: int):
    """C restatement (oracle/voxelize_ref.c).  points: (N, >=3) fp32."""
    lib = load_lib()
    pts = _f32(points)
    if pts.ndim != 2 or pts.shape[1] < 3:
        raise ValueError("points must be (N, >=3)")
    n, stride = pts.shape
    vs, mn, mx = _f32(grid.voxel_size), _f32(grid.range_min), _f32(grid.range_max)
    point_hash = np.empty(max(n, 1), np.int64)
    coords = np.empty((max(max_voxels, 1), 3), np.int32)
    indices = np.empty(max(n, 1), np.int64)
    splits = np.empty(max_voxels + 1, np.int64)
    counts = np.zeros(6, np.int64)
    rc = lib.p3p_oracle_voxelize(_ptr(pts), n, stride, _ptr(vs), _ptr(mn), _ptr(mx),
                                 grid.max_num_points, max_voxels, int(grid.drop_overflow), _ptr(point_hash), _ptr(coords),
                                 _ptr(indices), _ptr(splits), _ptr(counts))
    if rc != 0:
        raise RuntimeError("oracle voxelize failed")
    v, k = int(counts[0]), int(counts[1])
    return dict(voxel_coords=coords[:v].copy(), voxel_point_indices=indices[:k].copy(),
                voxel_point_row_splits=splits[:v + 1].copy(), point_hash=point_hash[:n].copy(),
                num_cells=int(counts[2]), extents=tuple(int(c) for c in counts[3:6]))


def voxelize_numpy(points: np.ndarray, grid: GridSpec, max_voxels: int):
    """Independent numpy restatement of A.1 (cross-checks the C one)."""
    pts = _f32(points)[:, :3]
    n = pts.shape[0]
    vs, mn, mx = _f32(grid.voxel_size), _f32(grid.range_min), _f32(grid.range_max)
    inv = (np.float32(1.0) / vs).astype(np.float32)
    extents = np.ceil(((mx - mn).astype(np.float32) * inv).astype(np.float32)).astype(np.int32).astype(np.int64)
    strides = np.array([1, extents[0], extents[0] * extents[1]], np.int64)
    num_cells = int(strides[2] * extents[2])
    invalid = np.iinfo(np.int64).max
    with np.errstate(invalid="ignore"):
        valid = np.all((pts >= mn) & (pts <= mx), axis=1)
        scaled = ((pts - mn).astype(np.float32) * inv).astype(np.float32)
    cells = np.zeros((n, 3), np.int64)
    cells[valid] = np.trunc(scaled[valid]).astype(np.int64)
    h = np.where(valid, cells @ strides, invalid).astype(np.int64)
    if grid.drop_overflow:
        h = np.where(h >= num_cells, invalid, h)
    order = np.lexsort((np.arange(n), h))  # ascending hash, ties by original index
    hs = h[order]
    n_valid = int(np.searchsorted(hs, invalid, side="left"))
    hs, order = hs[:n_valid], order[:n_valid]
    if n_valid:
        starts = np.flatnonzero(np.r_[True, hs[1:] != hs[:-1]])
    else:
        starts = np.zeros(0, np.int64)
    ends = np.r_[starts[1:], n_valid]
    starts, ends = starts[:max_voxels], ends[:max_voxels]
    keep = np.minimum(ends - starts, grid.max_num_points)
    idx = [order[s:s + k] for s, k in zip(starts, keep)]
    indices = np.concatenate(idx).astype(np.int64) if idx else np.zeros(0, np.int64)
    splits = np.r_[0, np.cumsum(keep)].astype(np.int64)
    coords = cells[order[starts]].astype(np.int32) if len(starts) else np.zeros((0, 3), np.int32)
    return dict(voxel_coords=coords, voxel_point_indices=indices, voxel_point_row_splits=splits,
                point_hash=np.where(h == invalid, -1, h), num_cells=num_cells, extents=tuple(int(e) for e in extents))


# --------------------------------------------------------------------------
# A.2  PointPillarsVoxelization.forward + PointPillars.voxelize
# --------------------------------------------------------------------------
def voxelization_forward(points: np.ndarray, grid: GridSpec, training: bool = False, impl=voxelize_c):
    """One tile -> (voxels (v,M,3) f32, coords (v,3) i32 [z,y,x], num_points (v,) i64, dense_idx (v,M) i64)."""
    pts = _f32(points)
    max_voxels = grid.max_voxels[0] if training else grid.max_voxels[1]
    ans = impl(pts, grid, max_voxels)
    M = grid.max_num_points
    splits, indices = ans["voxel_point_row_splits"], ans["voxel_point_indices"]
    v = len(splits) - 1
    # ragged_to_dense(indices, row_splits, M, default=-1) + 1 ; feats = cat([zeros(1,3), points])
    dense = np.full((v, M), -1, np.int64)
    for r in range(v):
        k = splits[r + 1] - splits[r]
        dense[r, :k] = indices[splits[r]:splits[r + 1]]
    feats = np.concatenate([np.zeros((1, pts.shape[1]), np.float32), pts], 0)
    out_voxels = feats[dense + 1]
    out_coords = ans["voxel_coords"][:, [2, 1, 0]]
    out_num = (splits[1:] - splits[:-1]).astype(np.int64)
    vs, mn, mx = _f32(grid.voxel_size), _f32(grid.range_min), _f32(grid.range_max)
    num_voxels = ((mx - mn).astype(np.float32) / vs).astype(np.float32).astype(np.int32)
    in_bounds = (out_coords[:, 1] < num_voxels[1]) & (out_coords[:, 2] < num_voxels[0])
    return (np.ascontiguousarray(out_voxels[in_bounds][:, :, :3]), np.ascontiguousarray(out_coords[in_bounds]),
            out_num[in_bounds], dense[in_bounds])


def batch_voxelize(tiles: Sequence[np.ndarray], grid: GridSpec, training: bool = False, impl=voxelize_c):
    """PointPillars.voxelize: loop over samples, concat, prepend batch id -> coors [b,z,y,x]."""
    voxels, coors, nums, dense = [], [], [], []
    for b, pts in enumerate(tiles):
        v, c, n, d = voxelization_forward(np.asarray(pts), grid, training, impl)
        voxels.append(v)
        coors.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
        nums.append(n)
        dense.append(d)
    M = grid.max_num_points
    return (np.concatenate(voxels, 0) if voxels else np.zeros((0, M, 3), np.float32),
            np.concatenate(nums, 0) if nums else np.zeros(0, np.int64),
            np.concatenate(coors, 0) if coors else np.zeros((0, 4), np.int32),
            np.concatenate(dense, 0) if dense else np.zeros((0, M), np.int64))


# --------------------------------------------------------------------------
# A.3 / A.4  PillarFeatureNet + PFNLayer (module + parameter names as upstream)
# --------------------------------------------------------------------------
class PFNLayer(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, last_layer: bool = False):
        super().__init__()
        self.last_vfe = last_layer
        self.units = out_channels if last_layer else out_channels // 2
        self.norm = nn.BatchNorm1d(self.units, eps=1e-3, momentum=0.01)
        self.linear = nn.Linear(in_channels, self.units, bias=False)

    def forward(self, inputs):
        x = self.linear(inputs)
        x = self.norm(x.permute(0, 2, 1).contiguous()).permute(0, 2, 1).contiguous()
        x = F.relu(x)
        x_max = torch.max(x, dim=1, keepdim=True)[0]  # over all M slots, padded ones included
        if self.last_vfe:
            return x_max
        return torch.cat([x, x_max.repeat(1, inputs.shape[1], 1)], dim=2)


class PillarFeatureNet(nn.Module):
    def __init__(self, in_channels=3, feat_channels=(64, 384), voxel_size=(8.0, 8.0, 100.0),
                 point_cloud_range=(0, 0, 0, 224, 224, 100), center_alias: bool = True):
        super().__init__()
        in_channels = in_channels + 5  # +3 cluster offsets, +2 pillar-centre offsets
        chans = [in_channels] + list(feat_channels)
        self.pfn_layers = nn.ModuleList(
            [PFNLayer(chans[i], chans[i + 1], last_layer=(i == len(chans) - 2)) for i in range(len(chans) - 1)])
        self.vx, self.vy = float(voxel_size[0]), float(voxel_size[1])
        self.x_offset = self.vx / 2 + point_cloud_range[0]
        self.y_offset = self.vy / 2 + point_cloud_range[1]
        self.center_alias = center_alias  # f_center is a *view* of features[:, :, :2]

    def decorate(self, features, num_points, coors):
        features = features.clone()  # the reference mutates its private gathered copy; never the caller's
        points_mean = features[:, :, :3].sum(dim=1, keepdim=True) / num_points.type_as(features).view(-1, 1, 1)
        f_cluster = features[:, :, :3] - points_mean
        f_center = features[:, :, :2] if self.center_alias else features[:, :, :2].clone()
        f_center[:, :, 0] = f_center[:, :, 0] - (coors[:, 3].to(features.dtype).unsqueeze(1) * self.vx + self.x_offset)
        f_center[:, :, 1] = f_center[:, :, 1] - (coors[:, 2].to(features.dtype).unsqueeze(1) * self.vy + self.y_offset)
        out = torch.cat([features, f_cluster, f_center], dim=-1)
        M = out.shape[1]
        mask = (torch.arange(M).view(1, -1) < num_points.view(-1, 1)).unsqueeze(-1).type_as(out)
        return out * mask

    def forward(self, features, num_points, coors):
        x = self.decorate(features, num_points, coors)
        for pfn in self.pfn_layers:
            x = pfn(x)
        return x.squeeze(dim=1)


# --------------------------------------------------------------------------
# A.5  PointPillarsScatter
# --------------------------------------------------------------------------
def scatter_canvas(voxel_features: torch.Tensor, coors: torch.Tensor, batch_size: int, channels: int, ny: int, nx: int):
    out = []
    for b in range(batch_size):
        canvas = torch.zeros(channels, nx * ny, dtype=voxel_features.dtype)
        m = coors[:, 0] == b
        this = coors[m]
        idx = (this[:, 2] * nx + this[:, 3]).long()
        vox = voxel_features[m].t()
        # `canvas[:, idx] = vox` with duplicate idx: CPU index_put_ is sequential, the later row
        # (higher hash) wins -- made explicit here so that it does not depend on ATen internals.
        for j in range(idx.numel()):
            canvas[:, idx[j]] = vox[:, j]
        out.append(canvas)
    if not out:
        return torch.zeros(0, channels, ny, nx)
    return torch.stack(out, 0).view(batch_size, channels, ny, nx)


# --------------------------------------------------------------------------
# the reference module, restated end to end
# --------------------------------------------------------------------------
class OraclePointPillarsEncoder(nn.Module):
    """CPU restatement of PointPillarsEncoder (pointpillars_o3d.py:11-107); same state_dict keys."""

    def __init__(self, grid: Optional[GridSpec] = None, center_alias: bool = True):
        super().__init__()
        self.grid = grid or GridSpec()
        g = self.grid
        self.voxel_encoder = PillarFeatureNet(
            in_channels=g.in_channels, feat_channels=g.feat_channels, voxel_size=g.voxel_size,
            point_cloud_range=list(g.range_min) + list(g.range_max), center_alias=center_alias)
        self.ny, self.nx = int(g.output_shape[0]), int(g.output_shape[1])
        self.channels = int(g.feat_channels[-1])

    @staticmethod
    def _tiles(x_lidar) -> List[np.ndarray]:
        if isinstance(x_lidar, torch.Tensor):
            if x_lidar.is_nested:
                return [t.detach().cpu().numpy() for t in x_lidar.unbind()]
            return [t.detach().cpu().numpy() for t in x_lidar]
        return [np.asarray(t) for t in x_lidar]

    @torch.no_grad()
    def voxelize(self, x_lidar):
        v, n, c, d = batch_voxelize(self._tiles(x_lidar), self.grid, self.training)
        return torch.from_numpy(v), torch.from_numpy(n), torch.from_numpy(c), torch.from_numpy(d)

    def pillar_features(self, x_lidar):
        voxels, num_points, coors, _ = self.voxelize(x_lidar)
        if voxels.shape[0] == 0:
            return torch.zeros(0, self.channels), coors, num_points
        return self.voxel_encoder(voxels, num_points, coors), coors, num_points

    def forward(self, x_lidar, return_flattened: bool = True):
        tiles = self._tiles(x_lidar)
        feats, coors, _ = self.pillar_features(tiles)
        x = scatter_canvas(feats, coors, len(tiles), self.channels, self.ny, self.nx)
        if return_flattened:
            return x.flatten(2).transpose(1, 2)
        return x


class OraclePatchEmbed(nn.Module):
    """timm PatchEmbed with flatten=False as used at early_fusion_vit.py:69-70,99: Conv2d(3,C,k=P,s=P,bias)."""

    def __init__(self, in_chans=3, embed_dim=384, patch=8):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch, stride=patch, bias=True)

    def forward(self, x):
        return self.proj(x)


def early_fusion_front(image_embed: OraclePatchEmbed, lidar_embed: OraclePointPillarsEncoder, x_image, x_lidar,
                       apply_lidar_dropout: bool = False):
    """early_fusion_vit.py:99-121 up to and including the concat (image channels first)."""
    xi = image_embed(x_image)
    xl = lidar_embed(x_lidar, return_flattened=False)
    if apply_lidar_dropout:
        xl = xl * 0.0
    return torch.cat((xi, xl), dim=1)


# --------------------------------------------------------------------------
# seeded synthetic inputs live in tools/synth.py (neutral module: bench.py's product arm uses them without
# importing the oracle); re-exported here for the tests
# --------------------------------------------------------------------------
import sys as _sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)
from tools.synth import synth_tile, synth_weights  # noqa: E402,F401


def vit_tokens(x_tokens: torch.Tensor, cls_token: torch.Tensor, pos_embed: torch.Tensor) -> torch.Tensor:
    """timm `VisionTransformer._pos_embed` in eval mode for the reference's `vit_small_patch8_224.dino` (one class
    token, no register tokens, `no_embed_class = False`, `pos_drop` = identity; the caller is
    R:pixelspointspolygons/models/pointpillars/pointpillars_vit.py:74 `self.vit.forward_features(x_lidar)`):
    x = cat([cls_token.expand(B, -1, -1), x], dim=1) + pos_embed.   x_tokens: (B, L, C) -> (B, 1 + L, C)."""
    B = x_tokens.shape[0]
    x = torch.cat([cls_token.reshape(1, 1, -1).expand(B, -1, -1).to(x_tokens.dtype), x_tokens], dim=1)
    return x + pos_embed.reshape(1, x.shape[1], -1).to(x_tokens.dtype)


def las_points_to_pixels(X, Y, Z, scales, offsets, top_left=None, height=224, width=224, res=0.25, z_hi=100.0,
                         variant="dataset") -> np.ndarray:
    """Restatement of the reference's LiDAR loaders on raw LAS integers (test infrastructure; numpy float64 exactly as
    the reference computes): `variant="dataset"` = P3Dataset.load_lidar_points (R:.../datasets/p3_coco.py:74-101),
    `variant="predict"` = Predictor.load_lidar_from_file (R:.../predict/predictor.py:116-137).  laspy's `las.x` is
    `las.X * scale + offset` in float64; sklearn's MinMaxScaler(feature_range=(0, z_hi)).fit_transform is
    `z * scale_ + min_` with `scale_ = (z_hi - 0) / handle_zeros(zmax - zmin)`, `min_ = 0 - zmin * scale_`
    (scikit-learn, `_handle_zeros_in_scale`: ranges below 10 eps become 1)."""
    pts = np.vstack((np.asarray(X, np.int32) * np.float64(scales[0]) + np.float64(offsets[0]),
                     np.asarray(Y, np.int32) * np.float64(scales[1]) + np.float64(offsets[1]),
                     np.asarray(Z, np.int32) * np.float64(scales[2]) + np.float64(offsets[2]))).transpose()
    if variant == "dataset":
        pts[:, :2] = (pts[:, :2] - np.asarray(top_left, np.float64)) / res
    else:
        pts[:, :2] = (pts[:, :2] - np.min(pts, axis=0)[:2]) / res
    pts[:, 1] = height - pts[:, 1]
    z = pts[:, -1]
    zmin, zmax = np.nanmin(z), np.nanmax(z)
    rng = zmax - zmin
    if rng < 10 * np.finfo(np.float64).eps:
        rng = 1.0
    scale_ = (z_hi - 0) / rng
    min_ = 0 - zmin * scale_
    z = z * scale_
    z = z + min_
    pts[:, -1] = z
    pts = pts.astype(np.float32)
    if variant == "dataset":
        pts[:, 0] = np.clip(pts[:, 0], 0, width)
        pts[:, 1] = np.clip(pts[:, 1], 0, height)
    return pts


def apply_d4_to_lidar(lidar: np.ndarray, group_element: str, center=(112, 112)) -> np.ndarray:
    """P3Dataset.apply_d4_augmentations_to_lidar (R:.../datasets/p3_coco.py:114-160) for an applied D4 transform:
    float32 in-place arithmetic about the tile centre, statement by statement as the reference."""
    lidar = np.array(lidar, dtype=np.float32, copy=True)
    lidar[:, :2] -= center
    if group_element == 'e':
        pass
    elif group_element == 'r90':
        lidar[:, [0, 1]] = lidar[:, [1, 0]]
        lidar[:, 1] = -lidar[:, 1]
    elif group_element == 'r180':
        lidar[:, 0] = -lidar[:, 0]
        lidar[:, 1] = -lidar[:, 1]
    elif group_element == 'r270':
        lidar[:, [0, 1]] = lidar[:, [1, 0]]
        lidar[:, 0] = -lidar[:, 0]
    elif group_element == 'v':
        lidar[:, 1] = -lidar[:, 1]
    elif group_element == 'hvt':
        lidar[:, [0, 1]] = lidar[:, [1, 0]]
        lidar[:, 0] = -lidar[:, 0]
        lidar[:, 1] = -lidar[:, 1]
    elif group_element == 'h':
        lidar[:, 0] = -lidar[:, 0]
    elif group_element == 't':
        lidar[:, [0, 1]] = lidar[:, [1, 0]]
    else:
        raise ValueError(f"Unknown group element {group_element}")
    lidar[:, :2] += center
    return lidar

