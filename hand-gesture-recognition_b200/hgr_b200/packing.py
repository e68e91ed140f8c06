"""state_dict -> packed parameter block.

Runs once per weight update (not on the hot path): folds BatchNorm
(reference model/gelan.py:46,56, eval mode: y = (x - mean) / sqrt(var + eps) *
gamma + beta) into per-channel fp32 scale/shift, re-orders conv weights from
(Cout, Cin, kh, kw) to the K-major [Cout][kh][kw][Cin] bf16 layout the
implicit-GEMM kernel's weight tiles use, and rounds the ViT matrices to bf16.
"""
from __future__ import annotations

import torch

from . import _lib

BN_EPS = 1e-5
LN_EPS = 1e-5


def sincos_table(h: int, w: int, dim: int = 256, temperature: float = 10000.0) -> torch.Tensor:
    """The fixed 2-D position table of the reference (model/transformer.py:9-26).

    omega = 1 / temperature**k with the RAW integer k = 0..dim/4-1 (not k/(dim/4-1)):
    most columns are therefore constant.  Reproduced as written, in fp32, on CPU.
    """
    if dim % 4:
        raise ValueError("dimension must be divisible by 4")
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    omega = 1.0 / (temperature ** torch.arange(dim // 4, dtype=torch.float32))
    yy = ys.flatten()[:, None] * omega[None, :]
    xx = xs.flatten()[:, None] * omega[None, :]
    return torch.cat((xx.sin(), xx.cos(), yy.sin(), yy.cos()), dim=1).to(torch.float32)


def fold_bn(sd, prefix: str):
    g = sd[prefix + ".bn.weight"].float()
    b = sd[prefix + ".bn.bias"].float()
    m = sd[prefix + ".bn.running_mean"].float()
    v = sd[prefix + ".bn.running_var"].float()
    scale = g / torch.sqrt(v + BN_EPS)
    return scale, b - m * scale


def packed_tensors(sd, image_size: int, device) -> dict:
    """name -> tensor for every entry of the library's parameter layout."""
    sd = {k: v.detach().to(device) for k, v in sd.items()}
    out = {}
    conv_prefixes = sorted({k[: -len(".conv.weight")] for k in sd if k.endswith(".conv.weight")})
    for p in conv_prefixes:
        w = sd[p + ".conv.weight"].float()
        scale, shift = fold_bn(sd, p)
        if p == "encoder.conv1":
            wk = (w * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(64, 27)
            wp = torch.zeros(64, 32, device=device)
            wp[:, :27] = wk
            out[p + ".w"] = wp.to(torch.bfloat16)
            out[p + ".shift"] = shift
        else:
            out[p + ".w"] = w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
            out[p + ".scale"] = scale
            out[p + ".shift"] = shift
    out["proj.w"] = sd["proj.weight"].float().reshape(256, 512).to(torch.bfloat16)
    f = image_size // 16
    out["decoder.pos_embedding"] = sincos_table(f, f).to(device).to(torch.bfloat16)
    out["decoder.cls_token"] = sd["decoder.cls_token"].float().reshape(256)
    cls_bf = out["decoder.cls_token"].to(torch.bfloat16).float()
    out["decoder.cls_token.stats"] = torch.stack(
        [cls_bf.mean(), torch.rsqrt(cls_bf.var(unbiased=False) + LN_EPS)]).float()

    def fold_ln(w, gamma, beta, bias=None):
        """LayerNorm folded into the Linear that consumes it: (W' = gamma (.) W in bf16, c = W' 1, d = W beta + b)."""
        w, gamma, beta = w.float(), gamma.float(), beta.float()
        wp = (w * gamma[None, :]).to(torch.bfloat16)
        d = w @ beta
        if bias is not None:
            d = d + bias.float()
        return wp, wp.float().sum(dim=1), d

    for l in range(4):
        a = f"decoder.transformer.layers.{l}.0."
        f = f"decoder.transformer.layers.{l}.1.net."
        out[a + "to_qkv.w"], out[a + "to_qkv.c"], out[a + "to_qkv.d"] = fold_ln(
            sd[a + "to_qkv.weight"], sd[a + "norm.weight"], sd[a + "norm.bias"])
        out[a + "to_out.w"] = sd[a + "to_out.weight"].float().to(torch.bfloat16)
        out[f + "1.w"], out[f + "1.c"], out[f + "1.d"] = fold_ln(
            sd[f + "1.weight"], sd[f + "0.weight"], sd[f + "0.bias"], sd[f + "1.bias"])
        out[f + "4.w"] = sd[f + "4.weight"].float().to(torch.bfloat16)
        out[f + "4.bias"] = sd[f + "4.bias"].float()
    for k in ("decoder.mlp_head.0.weight", "decoder.mlp_head.0.bias", "decoder.mlp_head.1.weight",
              "decoder.mlp_head.1.bias", "decoder.simple_decoder.1.bias"):
        out[k] = sd[k].float()
    v = sd["decoder.simple_decoder.1.weight"]
    out["decoder.simple_decoder.1.w"] = v.float().reshape(v.shape[0], 256).to(torch.bfloat16)
    return out


def pack(sd, image_size: int, num_joints: int, num_classes: int, device) -> torch.Tensor:
    """Returns the uint8 device tensor hgr_plan_create binds to."""
    layout = _lib.param_layout(image_size, num_joints, num_classes)
    total = _lib.load().hgr_param_bytes(image_size, num_joints, num_classes)
    tensors = packed_tensors(sd, image_size, device)
    block = torch.zeros(total, dtype=torch.uint8, device=device)
    for name, off, nbytes, dt, dims in layout:
        if name not in tensors:
            raise KeyError(f"state_dict has nothing for packed entry '{name}'")
        t = tensors[name].contiguous()
        want = torch.float32 if dt == _lib.F32 else torch.bfloat16
        n = 1
        for d in dims:
            n *= d
        if t.dtype != want or t.numel() != n:
            raise ValueError(f"packed entry '{name}': got {tuple(t.shape)} {t.dtype}, want {dims} {want}")
        block[off: off + nbytes].view(want).copy_(t.reshape(-1))
    return block
