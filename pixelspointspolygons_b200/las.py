"""LiDAR input front end: raw LAS integer coordinates -> pixel-space (N, 3) fp32 points, on the GPU.

Mirrors the numpy / scikit-learn body of `P3Dataset.load_lidar_points`
(R:pixelspointspolygons/datasets/p3_coco.py:74-101, `variant="dataset"`) and `Predictor.load_lidar_from_file`
(R:pixelspointspolygons/predict/predictor.py:116-137, `variant="predict"`), bit for bit, for a whole jagged batch:
the loader hands over `las.X / las.Y / las.Z` (int32) and the header's scales / offsets instead of `las.x / y / z`,
and gets the `values` tensor of the jagged batch `PointPillarsEncoder` consumes (same offsets).  SURVEY 8a row a1, 8f-3.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib


def tile_meta_bytes(tiles: Sequence[dict], variant: str) -> bytes:
    """Per-tile constants as the bytes of a p3p_las_tile array (what `HostPipeline.set_tiles` puts into a staging slot)."""
    if variant not in ("dataset", "predict"):
        raise ValueError("variant must be 'dataset' or 'predict'")
    arr = (_lib.LasTile * max(len(tiles), 1))()
    for i, t in enumerate(tiles):
        e = arr[i]
        for k in range(3):
            e.scale[k] = float(t["scales"][k])
            e.offset[k] = float(t["offsets"][k])
            if not e.scale[k] > 0.0:  # the tile extremes are taken over the integers: the int -> float64 map must be increasing
                raise ValueError(f"tile {i}: LAS header scale {e.scale[k]} must be positive")
        if variant == "dataset":
            e.left, e.top = float(t["top_left"][0]), float(t["top_left"][1])
            e.res, e.height, e.width = float(t.get("res_x", 0.25)), float(t["height"]), float(t["width"])
            e.origin_from_min, e.clip = 0, 1
        else:
            e.left = e.top = 0.0
            e.res, e.height, e.width = float(t.get("res_x", 0.25)), float(t.get("height", 224)), float(t.get("width", 224))
            e.origin_from_min, e.clip = 1, 0
        # replayed D4 element of the training augmentation (p3_coco.py:114-160): `d4` = group element name, centre =
        # (in_width // 2, in_height // 2)
        d4 = t.get("d4")
        e.d4 = _lib.P3P_D4[d4 if d4 is None else str(d4)]
        centre = t.get("center", (int(e.width) // 2, int(e.height) // 2))
        e.center_x, e.center_y = float(centre[0]), float(centre[1])
    return bytes(arr)[:C.sizeof(_lib.LasTile) * len(tiles)]


def _tile_meta(tiles: Sequence[dict], variant: str, dev) -> torch.Tensor:
    """Per-tile constants (p3p_las_tile) as a device byte tensor."""
    raw = tile_meta_bytes(tiles, variant)
    return torch.frombuffer(bytearray(raw if raw else b"\0" * C.sizeof(_lib.LasTile)), dtype=torch.uint8).to(dev)


def las_to_pixels(X: torch.Tensor, Y: torch.Tensor, Z: torch.Tensor, offsets: torch.Tensor, tiles: Sequence[dict],
                  z_hi: float = 100.0, variant: str = "dataset") -> torch.Tensor:
    """X, Y, Z: (ΣN) int32 CUDA tensors (the tiles' raw LAS coordinates, concatenated); offsets: (B + 1) int64 CUDA
    tensor; tiles: per tile a dict with `scales` (3), `offsets` (3) (las.header) and, for the dataset variant,
    `top_left` (2), `height`, `width`, optional `res_x` (default 0.25) and optional `d4` (the replayed group element
    'e', 'r90', 'r180', 'r270', 'v', 'hvt', 'h', 't' of the training augmentation when it was applied; absent / None: not
    applied); the predict variant uses `height` =
    `width` = 224 and res 0.25 unless given.  Returns (ΣN, 3) float32 on the same device."""
    if not (X.is_cuda and Y.is_cuda and Z.is_cuda and offsets.is_cuda):
        raise RuntimeError("las_to_pixels runs on CUDA only (sm_100a); no CPU fallback")
    if X.dtype != torch.int32 or Y.dtype != torch.int32 or Z.dtype != torch.int32 or offsets.dtype != torch.int64:
        raise TypeError("X, Y, Z must be int32 and offsets int64")
    B, total = offsets.numel() - 1, X.numel()
    if len(tiles) != B or Y.numel() != total or Z.numel() != total:
        raise ValueError("tiles / offsets / coordinate sizes disagree")
    dev = X.device
    meta = _tile_meta(tiles, variant, dev)
    out = torch.empty(total, 3, dtype=torch.float32, device=dev)
    mm = torch.empty(4 * max(B, 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().p3p_las_to_pixels(X.contiguous().data_ptr(), Y.contiguous().data_ptr(), Z.contiguous().data_ptr(),
                                          offsets.contiguous().data_ptr(), B, total, meta.data_ptr(), C.c_double(float(z_hi)),
                                          mm.data_ptr(), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "p3p_las_to_pixels")
    return out


def pack_las(X, Y, Z):
    """Host side of the packed transfer format for ONE tile: numpy int32 arrays (las.X, las.Y, las.Z) -> ((N, 3) uint16
    deltas, (3,) int32 base).  Raises when the tile spans 65536 steps or more of the LAS scale on some axis (send int32)."""
    import numpy as np

    xyz = np.stack([np.asarray(X), np.asarray(Y), np.asarray(Z)], axis=1).astype(np.int64)
    base = xyz.min(axis=0) if len(xyz) else np.zeros(3, np.int64)
    d = xyz - base
    if len(xyz) and d.max() > 65535:
        raise ValueError("tile spans more than 65535 steps of the LAS scale: use las_to_pixels with int32 coordinates")
    return d.astype(np.uint16), base.astype(np.int32)


class LasPackedFrontEnd:
    """las_packed_to_pixels with everything that does not change between batches of one shape prepared once (tile
    constants on the device, scratch, output buffer): what a serving loop calls per batch, and what a CUDA graph can hold."""

    def __init__(self, tiles: Sequence[dict], total: int, device, z_hi: float = 100.0, variant: str = "dataset", B: int = None):
        # (B without tiles: the per-tile constants arrive with every call, `meta=` -- HostPipeline keeps them in its slots)
        self.B, self.total, self.device, self.z_hi = (len(tiles) if B is None else int(B)), int(total), device, float(z_hi)
        self.meta = _tile_meta(tiles, variant, device) if len(tiles) else None
        self.mm = torch.empty(4 * max(self.B, 1), dtype=torch.int32, device=device)
        self.out = torch.empty(self.total, 3, dtype=torch.float32, device=device)

    def __call__(self, deltas: torch.Tensor, base: torch.Tensor, offsets: torch.Tensor, out: torch.Tensor = None,
                 meta: torch.Tensor = None) -> torch.Tensor:
        if deltas.dtype != torch.uint16 or base.dtype != torch.int32 or offsets.dtype != torch.int64:
            raise TypeError("deltas must be uint16 (N, 3), base int32 (B, 3), offsets int64")
        if not (deltas.is_cuda and base.is_cuda and offsets.is_cuda):
            raise RuntimeError("las_packed_to_pixels runs on CUDA only (sm_100a); no CPU fallback")
        if deltas.numel() != 3 * self.total or base.numel() != 3 * self.B or offsets.numel() != self.B + 1:
            raise ValueError("deltas / base / offsets sizes disagree with the prepared batch shape")
        out = self.out if out is None else out
        meta = self.meta if meta is None else meta
        if meta is None or not meta.is_cuda or meta.numel() * meta.element_size() < self.B * C.sizeof(_lib.LasTile):
            raise ValueError("meta must be a CUDA byte tensor holding one p3p_las_tile per tile")
        with torch.cuda.device(self.device):
            rc = _lib.lib().p3p_las_packed_to_pixels(deltas.data_ptr(), base.data_ptr(), offsets.data_ptr(), self.B, self.total,
                                                     meta.data_ptr(), C.c_double(self.z_hi), self.mm.data_ptr(), out.data_ptr(),
                                                     torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "p3p_las_packed_to_pixels")
        return out


def las_packed_to_pixels(deltas: torch.Tensor, base: torch.Tensor, offsets: torch.Tensor, tiles: Sequence[dict],
                         z_hi: float = 100.0, variant: str = "dataset") -> torch.Tensor:
    """las_to_pixels fed with the packed transfer format: deltas (ΣN, 3) uint16 and base (B, 3) int32 CUDA tensors
    (X = base + delta; `pack_las` makes them per tile on the host) -- 6 bytes per point over PCIe instead of 12.
    Bit-identical to las_to_pixels on the same integers."""
    fe = LasPackedFrontEnd(tiles, deltas.shape[0], deltas.device, z_hi, variant)
    return fe(deltas.contiguous(), base.contiguous(), offsets.contiguous())
