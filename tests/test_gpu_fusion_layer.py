"""GPU parity of SURVEY 8f rank 1 / 5: the reference's `fusion_layer` (Conv2d 3x3 + BatchNorm2d + ReLU, flatten,
transpose; early_fusion_vit.py:75-79,123) and `proj` tails (bilinear Upsample + Conv2d 3x3 + BatchNorm2d + ReLU;
pointpillars_vit_cnn.py:20-25,36, early_fusion_vit_cnn.py:78-83,102) through libp3p.so, against the REAL torch modules on
the CPU in fp32 (a pinned oracle: torch is the reference's own implementation of these layers).
Bars: 1e-3 of scale with fp16 operands (fp32 contract), 1e-2 with bf16 operands, fp32 accumulation in both."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import pillars_oracle as po
from pixelspointspolygons_b200 import default_cfg
from pixelspointspolygons_b200.fusion import ConvBnRelu3x3, EarlyFusionFrontEnd, ProjTail
from test_gpu_parity import assert_close, to_nested

pytestmark = pytest.mark.gpu

TOL = {"fp16": 1e-3, "bf16": 1e-2}


def randomize_bn(bn: nn.BatchNorm2d, g):
    with torch.no_grad():
        bn.weight.copy_(torch.rand(bn.num_features, generator=g) + 0.5)
        bn.weight[::5] *= -1.0
        bn.bias.copy_(torch.randn(bn.num_features, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(bn.num_features, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(bn.num_features, generator=g) + 0.5)


def torch_fusion_layer(cin, cout, seed):
    g = torch.Generator().manual_seed(seed)
    m = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)).eval()
    with torch.no_grad():
        m[0].weight.copy_(torch.randn(m[0].weight.shape, generator=g) / (9 * cin) ** 0.5)
        m[0].bias.copy_(torch.randn(cout, generator=g) * 0.1)
    randomize_bn(m[1], g)
    return m


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 768, 384, 28, 28), (1, 128, 200, 12, 20), (3, 64, 64, 8, 16)])
def test_conv_bn_relu_matches_torch(cuda_device, prec, shape):
    B, cin, cout, H, W = shape
    ref = torch_fusion_layer(cin, cout, 5)
    mod = ConvBnRelu3x3(cin, cout, precision=prec).eval()
    mod.load_state_dict(ref.state_dict())  # same keys as the reference's nn.Sequential
    mod = mod.to(cuda_device)
    x = torch.randn(B, cin, H, W, generator=torch.Generator().manual_seed(1))
    x[:, :, 0, :] += 2.0  # make the borders matter (zero padding)
    with torch.no_grad():
        want = ref(x.clone())
        got = mod(x.to(cuda_device))
    assert got.shape == want.shape
    assert_close(got, want, TOL[prec], f"conv3x3 {shape} {prec}")
    # token rows: the flatten(2).transpose(1, 2) of the reference is the store address
    x16 = x.to(cuda_device).permute(0, 2, 3, 1).contiguous().to(mod.operand_dtype)
    tokens = torch.empty(B, H * W, cout, device=cuda_device)
    mod.forward_nhwc(x16, tokens, 1)
    torch.cuda.synchronize()
    assert_close(tokens, want.flatten(2).transpose(1, 2), TOL[prec], f"conv3x3 tokens {shape} {prec}")


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_fusion_tokens_match_reference_modules(cuda_device, prec):
    """EarlyFusionViT.forward up to the ViT: image_embed + lidar_embed + concat + fusion_layer + flatten + transpose."""
    cfg = default_cfg(device=str(cuda_device), p3p_precision=prec)
    fe = EarlyFusionFrontEnd(cfg).eval()
    sd, sdi = po.synth_weights(3)
    fe.lidar_embed.load_state_dict(sd)
    fe.image_embed.load_state_dict(sdi)
    ref_fl = torch_fusion_layer(768, 384, 9)
    fe.fusion_layer.load_state_dict(ref_fl.state_dict())
    fe = fe.to(cuda_device)
    tiles = [po.synth_tile(20000, 41), po.synth_tile(5000, 42, clustered=True), po.synth_tile(0, 43)]
    imgs = torch.rand(3, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    ref_enc = po.OraclePointPillarsEncoder(po.GridSpec()).eval()
    ref_enc.load_state_dict(sd)
    ref_pe = po.OraclePatchEmbed().eval()
    ref_pe.load_state_dict(sdi)
    with torch.no_grad():
        concat = po.early_fusion_front(ref_pe, ref_enc, imgs, tiles)
        want = ref_fl(concat).flatten(2).transpose(1, 2)
        got = fe.forward_tokens(imgs.to(cuda_device), to_nested(tiles, cuda_device))
        # LiDAR dropout: the LiDAR half of the convolution's input is zero
        concat0 = concat.clone()
        concat0[:, 384:] = 0.0
        want0 = ref_fl(concat0).flatten(2).transpose(1, 2)
        got0 = fe.forward_tokens(imgs.to(cuda_device), to_nested(tiles, cuda_device), lidar_zero=True)
    torch.cuda.synchronize()
    tol = 2e-3 if prec == "fp16" else 2e-2  # two 16-bit roundings in series (activations, then the convolution's operands)
    assert_close(got, want, tol, f"fusion tokens {prec}")
    assert_close(got0, want0, tol, f"fusion tokens, LiDAR dropout {prec}")


def test_proj_tail_matches_torch(cuda_device):
    """FFL / HiSup tail at its real size: tokens (B, 784, 384) -> upsample 224 x 224 -> conv 384 -> 256 + BN + ReLU."""
    g = torch.Generator().manual_seed(11)
    ref = nn.Sequential(nn.Upsample(size=224, mode="bilinear", align_corners=False), nn.Conv2d(384, 256, kernel_size=3, padding=1),
                        nn.BatchNorm2d(256), nn.ReLU(inplace=True)).eval()
    with torch.no_grad():
        ref[1].weight.copy_(torch.randn(ref[1].weight.shape, generator=g) / (9 * 384) ** 0.5)
        ref[1].bias.copy_(torch.randn(256, generator=g) * 0.1)
    randomize_bn(ref[2], g)
    mod = ProjTail(384, 256, 224).eval()
    mod.load_state_dict(ref.state_dict())
    mod = mod.to(cuda_device)
    tokens = torch.randn(1, 785, 384, generator=g)  # with the class token in row 0
    with torch.no_grad():
        x = tokens[:, 1:, :].permute(0, 2, 1).reshape(1, 384, 28, 28)
        want = ref(x)
        got = mod.forward_tokens(tokens.to(cuda_device), 28, 28, skip_rows=1)
        got2 = mod(x.to(cuda_device))
    torch.cuda.synchronize()
    assert_close(got, want, 1e-3, "proj tail (tokens in place)")
    assert_close(got2, want, 1e-3, "proj tail (NCHW drop-in)")


def test_proj_tail_small_and_bf16(cuda_device):
    g = torch.Generator().manual_seed(12)
    ref = nn.Sequential(nn.Upsample(size=(24, 32), mode="bilinear", align_corners=False), nn.Conv2d(64, 96, kernel_size=3, padding=1),
                        nn.BatchNorm2d(96), nn.ReLU(inplace=True)).eval()
    randomize_bn(ref[2], g)
    mod = ProjTail(64, 96, (24, 32), precision="bf16").eval()
    mod.load_state_dict(ref.state_dict())
    mod = mod.to(cuda_device)
    x = torch.randn(2, 64, 5, 7, generator=g)
    with torch.no_grad():
        want = ref(x)
        got = mod(x.to(cuda_device))
    torch.cuda.synchronize()
    assert_close(got, want, 1e-2, "proj tail small bf16")
