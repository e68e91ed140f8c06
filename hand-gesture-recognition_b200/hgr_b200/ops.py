"""Device versions of the reference's host-side helpers around the model.

get_max_preds         reference libs/utils.py:4-32
crop_normalize        reference detect.py:106-112 / libs/load.py:46-50
crop_warp_normalize   reference detect.py:92-117 (get_affine_transform + cv2.warpAffine + normalise)
pose_accuracy         reference libs/metrics.py:31-62 (PCK on the decoded keypoints)
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


def box_to_affine(bbox, size: int) -> np.ndarray:
    """The 2x3 float64 matrix detect.py:93-96 builds for a detector box (x1, y1, x2, y2): centre, side
    max(w, h), rotation 0, via libs/transforms.py:20-54 (three float32 point pairs; cv2.getAffineTransform solves
    the 6x6 system in double precision - so does numpy here)."""
    x1, y1, x2, y2 = (float(v) for v in bbox)
    c = np.array([(x1 + x2) / 2, (y1 + y2) / 2], dtype=np.float32)
    src_w = max(x2 - x1, y2 - y1) * 1.0
    dst_w = dst_h = float(size)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = c
    src[1, :] = c + np.array([0.0 * 1.0 - (src_w * -0.5) * 0.0, 0.0 * 0.0 + (src_w * -0.5) * 1.0])
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)
    for pts in (src, dst):
        d = pts[0] - pts[1]
        pts[2, :] = pts[1] + np.array([-d[1], d[0]], dtype=np.float32)
    a = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        a[2 * i] = [src[i, 0], src[i, 1], 1, 0, 0, 0]
        a[2 * i + 1] = [0, 0, 0, src[i, 0], src[i, 1], 1]
        b[2 * i], b[2 * i + 1] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(a, b).reshape(2, 3)


def invert_affine(trans) -> np.ndarray:
    """dst -> src map exactly as cv::warpAffine derives it from `trans` (float64, same operation order)."""
    m = np.array(trans, dtype=np.float64).reshape(6).copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def crop_warp_normalize(frames_u8: torch.Tensor, boxes, frame_index=None, size: int = 192, dtype=torch.float32,
                        trans=None) -> torch.Tensor:
    """detect.py:92-117 on the device: frames (F, Hf, Wf, 3) or (Hf, Wf, 3) uint8 CUDA, one detector box
    (x1, y1, x2, y2) per crop (or explicit 2x3 `trans` matrices) -> (N, 3, size, size) normalised crops,
    bit-identical to cv2.warpAffine(INTER_LINEAR) + the reference's normalisation."""
    if not isinstance(frames_u8, torch.Tensor) or not frames_u8.is_cuda or frames_u8.dtype != torch.uint8:
        raise RuntimeError("crop_warp_normalize runs on the GPU only: pass a CUDA uint8 tensor")
    fr = frames_u8 if frames_u8.dim() == 4 else frames_u8[None]
    if fr.shape[-1] != 3:
        raise TypeError("expected uint8 frames (..., Hf, Wf, 3)")
    fr = fr.contiguous()
    mats = [np.asarray(t, dtype=np.float64) for t in trans] if trans is not None else \
        [box_to_affine(b, size) for b in boxes]
    n = len(mats)
    idx = np.zeros(n, dtype=np.int32) if frame_index is None else np.asarray(frame_index, dtype=np.int32)
    if idx.shape != (n,) or (n and (idx.min() < 0 or idx.max() >= fr.shape[0])):
        raise ValueError("frame_index must hold one valid frame number per crop")
    inv = np.stack([invert_affine(m) for m in mats]) if n else np.zeros((0, 6))
    out = torch.empty(n, 3, size, size, dtype=dtype, device=fr.device)
    if n:
        d_inv = torch.from_numpy(np.ascontiguousarray(inv)).to(fr.device)
        d_idx = torch.from_numpy(idx).to(fr.device)
        with torch.cuda.device(fr.device):
            _lib.check(_lib.load().hgr_crop_warp_normalize(fr.data_ptr(), fr.shape[0], fr.shape[1], fr.shape[2],
                                                           d_idx.data_ptr(), d_inv.data_ptr(), n, size, out.data_ptr(),
                                                           _lib.F32 if dtype == torch.float32 else _lib.BF16,
                                                           _stream(fr.device)), "hgr_crop_warp_normalize")
        out.record_stream(torch.cuda.current_stream(fr.device))
    return out


def pose_accuracy(output: torch.Tensor, target: torch.Tensor, thr: float = 0.5):
    """libs.metrics.pose_accuracy on the device: predicted and ground-truth heatmaps (B, J, H, W) CUDA ->
    (acc (J + 1,) float64, avg_acc, cnt, pred (B, J, 2)) as CUDA tensors, without the device->host copy of the
    heatmaps that train.py:71-73 makes every step.  `float(avg_acc)`, `int(cnt)` give the reference's scalars."""
    if not (isinstance(output, torch.Tensor) and isinstance(target, torch.Tensor) and output.is_cuda and target.is_cuda):
        raise RuntimeError("pose_accuracy runs on the GPU only: pass CUDA tensors")
    if output.dim() != 4 or output.shape != target.shape:
        raise ValueError("output and target must be (B, J, H, W) heatmaps of the same shape")
    b, j, h, w = output.shape
    pred, _ = get_max_preds(output.detach())
    tgt, _ = get_max_preds(target.detach())
    dev = output.device
    counts = torch.empty(2 * j, dtype=torch.int32, device=dev)
    acc = torch.empty(j + 1, dtype=torch.float64, device=dev)
    avg_cnt = torch.empty(2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().hgr_pose_accuracy(pred.data_ptr(), tgt.data_ptr(), b, j, h, w, float(thr),
                                                 counts.data_ptr(), acc.data_ptr(), avg_cnt.data_ptr(), _stream(dev)),
                   "hgr_pose_accuracy")
    return acc, avg_cnt[0], avg_cnt[1].to(torch.int64), pred
