"""Device versions of the reference's host-side helpers around the model.

get_max_preds    reference libs/utils.py:4-32
crop_normalize   reference detect.py:106-112 / libs/load.py:46-50
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def get_max_preds(batch_heatmaps):
    """get predictions from score maps (B, J, H, W) -> (preds (B, J, 2), maxvals (B, J, 1)).

    torch CUDA tensor in -> torch CUDA tensors out (no device->host copy);
    numpy array in -> numpy arrays out, like the reference's signature.
    """
    is_np = isinstance(batch_heatmaps, np.ndarray)
    assert is_np or isinstance(batch_heatmaps, torch.Tensor), \
        'batch_heatmaps should be numpy.ndarray or a CUDA torch.Tensor'
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    if is_np:
        if not torch.cuda.is_available():
            raise RuntimeError("get_max_preds runs on the GPU only; no CUDA device is available")
        h = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, dtype=np.float32)).cuda()
    else:
        h = batch_heatmaps
        if not h.is_cuda:
            raise RuntimeError("get_max_preds runs on the GPU only: pass a CUDA tensor")
        if h.dtype not in (torch.float32, torch.bfloat16):
            h = h.float()
        h = h.contiguous()
    b, j, hh, ww = h.shape
    preds = torch.empty(b, j, 2, dtype=torch.float32, device=h.device)
    maxvals = torch.empty(b, j, 1, dtype=torch.float32, device=h.device)
    if b * j > 0:
        with torch.cuda.device(h.device):
            _lib.check(_lib.load().hgr_get_max_preds(h.data_ptr(), _lib.F32 if h.dtype == torch.float32 else _lib.BF16,
                                                     b, j, hh, ww, preds.data_ptr(), maxvals.data_ptr(),
                                                     _stream(h.device)), "hgr_get_max_preds")
    if is_np:
        return preds.cpu().numpy(), maxvals.cpu().numpy()
    return preds, maxvals


def crop_normalize(img_hwc_u8: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """(B, H, W, 3) or (H, W, 3) uint8 on CUDA -> (B, 3, H, W) normalised crop."""
    if not isinstance(img_hwc_u8, torch.Tensor) or not img_hwc_u8.is_cuda:
        raise RuntimeError("crop_normalize runs on the GPU only: pass a CUDA uint8 tensor")
    if img_hwc_u8.dtype != torch.uint8 or img_hwc_u8.shape[-1] != 3:
        raise TypeError("expected uint8 (..., H, W, 3)")
    x = img_hwc_u8 if img_hwc_u8.dim() == 4 else img_hwc_u8[None]
    x = x.contiguous()
    b, h, w, _ = x.shape
    out = torch.empty(b, 3, h, w, dtype=dtype, device=x.device)
    if b * h * w > 0:
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().hgr_crop_normalize(x.data_ptr(), out.data_ptr(),
                                                      _lib.F32 if dtype == torch.float32 else _lib.BF16, b, h, w,
                                                      _stream(x.device)), "hgr_crop_normalize")
    return out
