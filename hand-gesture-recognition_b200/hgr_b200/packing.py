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
    for k, v in sd.items():
        if not k.startswith("decoder."):
            continue
        if k.endswith("to_qkv.weight") or k.endswith("to_out.weight") or k.endswith("net.1.weight") \
                or k.endswith("net.4.weight"):
            out[k[: -len("weight")] + "w"] = v.float().to(torch.bfloat16)
        elif k == "decoder.simple_decoder.1.weight":
            out["decoder.simple_decoder.1.w"] = v.float().reshape(v.shape[0], 256).to(torch.bfloat16)
        elif k != "decoder.cls_token":
            out[k] = v.float()
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
