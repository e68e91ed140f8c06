"""Shared test utilities: error metrics and raw C-ABI call helpers."""
from __future__ import annotations

import torch

BF16_EPS = 2.0 ** -8  # half an ulp of bf16 relative to the value's binade is 2^-9; one ulp 2^-8


def rel_l2(out: torch.Tensor, ref: torch.Tensor) -> float:
    out, ref = out.double().flatten(), ref.double().flatten()
    return float((out - ref).norm() / ref.norm().clamp_min(1e-30))


def max_abs(out: torch.Tensor, ref: torch.Tensor) -> float:
    return float((out.double() - ref.double()).abs().max())


def report(name: str, out: torch.Tensor, ref: torch.Tensor) -> tuple[float, float]:
    """Prints and returns (rel-L2, max-abs / max|ref|)."""
    out, ref = out.detach().float().cpu(), ref.detach().float().cpu()
    assert out.shape == ref.shape, f"{name}: shape {tuple(out.shape)} vs {tuple(ref.shape)}"
    r, m = rel_l2(out, ref), max_abs(out, ref)
    scale = float(ref.abs().max())
    bad = (out - ref).abs().flatten().topk(min(3, out.numel()))
    print(f"[parity] {name}: rel_l2={r:.3e} max_abs={m:.3e} (max|ref|={scale:.3e}) nan_out={int(out.isnan().sum())} "
          f"worst_idx={[int(i) for i in bad.indices]}", flush=True)
    return r, m / max(scale, 1e-30)


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def nhwc_bf16(t_nchw: torch.Tensor, device) -> torch.Tensor:
    return t_nchw.permute(0, 2, 3, 1).contiguous().to(device=device, dtype=torch.bfloat16)


def nchw_f32(t_nhwc: torch.Tensor) -> torch.Tensor:
    return t_nhwc.float().permute(0, 3, 1, 2).contiguous().cpu()
