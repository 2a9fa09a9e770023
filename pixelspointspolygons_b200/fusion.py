"""Early-fusion front end: image patch embedding + LiDAR pillar encoder -> one (B, 2C, ny, nx) concat buffer.

Mirrors the part of the reference's fusion encoders that is on the hot path
(R:pixelspointspolygons/models/fusion_layers/early_fusion_vit.py:43-52,69-72,96-121 and
early_fusion_vit_cnn.py:41-50,67-70,90-92):

    x_image = self.image_embed(x_image)                          # timm PatchEmbed, flatten=False -> (B, C, ny, nx)
    x_lidar = self.lidar_embed(x_lidar, return_flattened=False)  # PointPillarsEncoder          -> (B, C, ny, nx)
    if cfg.experiment.lidar_dropout is not None and rand <= p:  x_lidar = x_lidar * 0.0
    x = torch.cat((x_image, x_lidar), dim=1)                     # image channels first

Submodule and parameter names are the reference's (`image_embed.proj.{weight,bias}`,
`lidar_embed.voxel_encoder.pfn_layers.*`), so a checkpoint's `encoder.image_embed.*` /
`encoder.lidar_embed.*` entries load unchanged.  In eval mode both halves are written by libp3p.so kernels directly
into the concat buffer (no intermediate tensors, no cat copy); everything after the concat (`fusion_layer`, the
ViT) is outside this package.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import P3P_DTYPE_BF16, P3P_DTYPE_F32, P3P_LAYOUT_NCHW, P3P_PRECISION
from .encoder import PointPillarsEncoder, _get


class PatchEmbed(nn.Module):
    """timm.layers.PatchEmbed as the reference uses it (`flatten = False`, norm = Identity, NCHW output)."""

    def __init__(self, img_size=224, patch_size=8, in_chans=3, embed_dim=384, bias=True, precision: str = "fp16"):
        super().__init__()
        self.img_size, self.patch_size, self.in_chans, self.embed_dim = int(img_size), int(patch_size), int(in_chans), int(embed_dim)
        self.flatten = False
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)
        self.norm = nn.Identity()
        self.precision = precision

    def forward_into(self, x: torch.Tensor, out: torch.Tensor, c_total: int, c_offset: int, precision: Optional[str] = None):
        if not x.is_cuda:
            raise RuntimeError("PatchEmbed runs on CUDA only (sm_100a); x_image is on " + str(x.device))
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != self.in_chans:
            raise TypeError(f"x_image must be float32 (B, {self.in_chans}, H, W)")
        x = x.contiguous()
        w = self.proj.weight.detach().contiguous()
        b = self.proj.bias.detach().contiguous() if self.proj.bias is not None else None
        B, _, H, W = x.shape
        dt = P3P_DTYPE_F32 if out.dtype == torch.float32 else P3P_DTYPE_BF16
        with torch.cuda.device(x.device):
            rc = _lib.lib().p3p_patch_embed(x.data_ptr(), B, self.in_chans, H, W, self.patch_size, w.data_ptr(),
                                            b.data_ptr() if b is not None else None, self.embed_dim,
                                            P3P_PRECISION[precision or self.precision], out.data_ptr(), dt, c_total, c_offset,
                                            torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(rc, "p3p_patch_embed")
        return out

    def forward(self, x):
        if self.training and torch.is_grad_enabled():
            y = self.proj(x)  # autograd route of the optional training step
        else:
            B, _, H, W = x.shape
            y = torch.empty(B, self.embed_dim, H // self.patch_size, W // self.patch_size, dtype=torch.float32, device=x.device)
            self.forward_into(x, y, self.embed_dim, 0)
        if self.flatten:
            y = y.flatten(2).transpose(1, 2)
        return self.norm(y)


class EarlyFusionFrontEnd(nn.Module):
    """`EarlyFusionViT.forward` / `EarlyFusionViTCNN.forward` up to and including the concat."""

    def __init__(self, cfg, local_rank: int = 0):
        super().__init__()
        self.cfg = cfg
        self.local_rank = local_rank
        verbosity = getattr(logging, str(_get(_get(cfg, "run_type"), "logging", "INFO")).upper(), logging.INFO)
        self.logger = logging.getLogger(f"{self.__class__.__name__}[{local_rank}]")
        self.logger.setLevel(verbosity)
        enc = cfg.experiment.encoder
        dim = int(enc.patch_feature_dim)
        # early_fusion_vit.py:43-52
        output_shape = [int(enc.patch_feature_width), int(enc.patch_feature_height)]
        self.lidar_embed = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, dim]},
                                               scatter={"in_channels": dim, "output_shape": output_shape}, local_rank=local_rank)
        # early_fusion_vit.py:69-70: the ViT's own PatchEmbed, re-parented, flatten switched off
        self.image_embed = PatchEmbed(img_size=int(_get(enc, "in_size", enc.in_height)), patch_size=int(_get(enc, "patch_size", 8)),
                                      in_chans=3, embed_dim=dim, precision=self.lidar_embed.precision)
        self.channels = dim
        self._side_streams = {}

    def _dropout_now(self) -> bool:
        """early_fusion_vit.py:113-119 (one draw per forward, whole batch)."""
        p = _get(_get(self.cfg, "experiment"), "lidar_dropout", None)
        if p is None:
            return False
        return bool(torch.rand(1).item() <= float(p))

    def forward_into(self, x_image, x_lidar, out: torch.Tensor, lidar_zero: Optional[bool] = None):
        """Eval-mode fused path: both halves written in place into `out` (B, 2C, ny, nx); returns `out`."""
        dim = self.channels
        lidar_zero = self._dropout_now() if lidar_zero is None else bool(lidar_zero)
        # the two halves are independent until the concat: the patch embedding runs on a side stream next to the
        # voxelizer + PFN (fork / join by events, so a CUDA graph captures them as parallel branches)
        dev = out.device
        cur = torch.cuda.current_stream(dev)
        side = self._side_streams.get(dev)
        if side is None:
            side = self._side_streams[dev] = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.image_embed.forward_into(x_image, out, 2 * dim, 0)
        self.lidar_embed.encode_into(x_lidar, out, P3P_LAYOUT_NCHW, c_total=2 * dim, c_offset=dim, lidar_zero=lidar_zero)
        cur.wait_stream(side)
        return out

    def forward(self, x_image, x_lidar):
        if self.training:  # BatchNorm batch statistics: the dense autograd route of the optional training step
            xi = self.image_embed(x_image)
            xl = self.lidar_embed(x_lidar, return_flattened=False)
            if self._dropout_now():
                xl = xl * 0.0
            return torch.cat((xi, xl), dim=1)
        B = x_image.shape[0]
        le = self.lidar_embed
        out = torch.empty(B, 2 * self.channels, le.ny, le.nx, dtype=le.out_dtype, device=x_image.device)
        return self.forward_into(x_image, x_lidar, out)
