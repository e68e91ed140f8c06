"""Batched, host-facing version of the classifier stage of the reference's inference app.

reference detect.py:92-117,140-155 does, per frame and per hand:
    uint8 BGR crop -> /255, mean/std, CHW -> classifier -> argmax(label), get_max_preds(heatmap)
This class does the same for a whole batch of crops that live in HOST memory:
pinned uint8 crops are copied to the device, normalised (crop_normalize
kernel), run through the MultiTaskNet plan with the keypoint decode fused into
the pose head's epilogue (hgr_forward_keypoints: the fp32 heatmaps - 198 MB per
1024 crops - are never written, the decode is bit-identical to get_max_preds on
them) and only the logits, keypoints and their confidences travel back -
19*4 + 21*3*4 bytes per hand instead of a 194 KB heatmap.

Two lanes (copy stream + buffers + plan each) are used alternately so that the
host->device copy of batch i+1 and the device->host copy of batch i-1 overlap
the kernels of batch i.  ALL kernels run on one compute stream: the persistent
GEMM kernels of two batches never interleave (each would otherwise take every
SM and double the tail effects), only the copies run beside them.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .model import MultiTaskNet, _Plan


class _Lane:
    def __init__(self, model: MultiTaskNet, batch: int, device, dtype):
        s = model.image_size[0]
        self.stream = torch.cuda.Stream(device)
        self.plan = _Plan(s, model.num_joints, model.num_classes, batch, model._packed_params(device))
        self.h_crops = torch.empty(batch, s, s, 3, dtype=torch.uint8).pin_memory()
        self.d_crops = torch.empty(batch, s, s, 3, dtype=torch.uint8, device=device)
        self.d_x = torch.empty(batch, 3, s, s, dtype=dtype, device=device)
        self.d_logits = torch.empty(batch, model.num_classes, dtype=torch.float32, device=device)
        # only when the fused decode is unavailable (image_size > 320 or HGR_POSE_TC=0) do heatmaps reach memory
        fused = _lib.load().hgr_keypoints_fused(s, model.num_joints) == 1
        self.d_heat = None if fused else torch.empty(batch, model.num_joints, s // 4, s // 4, dtype=torch.float32,
                                                     device=device)
        self.d_preds = torch.empty(batch, model.num_joints, 2, dtype=torch.float32, device=device)
        self.d_maxvals = torch.empty(batch, model.num_joints, 1, dtype=torch.float32, device=device)
        self.h_logits = torch.empty(batch, model.num_classes, dtype=torch.float32).pin_memory()
        self.h_preds = torch.empty(batch, model.num_joints, 2, dtype=torch.float32).pin_memory()
        self.h_maxvals = torch.empty(batch, model.num_joints, 1, dtype=torch.float32).pin_memory()
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.busy = False


class HandPipeline:
    def __init__(self, model: MultiTaskNet, batch: int, compute_dtype=torch.bfloat16, lanes: int = 2):
        if model.training:
            raise RuntimeError("HandPipeline needs model.eval()")
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("HandPipeline needs the model on a CUDA device; there is no CPU path")
        self.model, self.batch, self.device, self.dtype = model, batch, p.device, compute_dtype
        with torch.cuda.device(self.device):
            self.lanes = [_Lane(model, batch, self.device, compute_dtype) for _ in range(lanes)]
            self.compute = torch.cuda.Stream(self.device)
        self._next = 0
        self._params = self.lanes[0].plan.params
        self.h2d_bytes = self.lanes[0].h_crops.numel()
        self.d2h_bytes = 4 * (self.lanes[0].h_logits.numel() + self.lanes[0].h_preds.numel()
                              + self.lanes[0].h_maxvals.numel())
        # crop_normalize + the plan's launches (the keypoint decode is part of the last one)
        self.launches_per_batch = self.lanes[0].plan.launches() + 1

    def _refresh_plans(self):
        """The lanes' plans point into the weights packed when they were built.  After load_state_dict / an optimiser
        step / .to() the model re-packs: rebuild the plans against the new block (the old block stays alive until then,
        each plan holds a reference to it)."""
        params = self.model._packed_params(self.device)
        if params is not self._params:
            if any(ln.busy for ln in self.lanes):
                raise RuntimeError("the model's parameters changed while a batch is in flight; collect() it first")
            m = self.model
            with torch.cuda.device(self.device):
                torch.cuda.current_stream(self.device).synchronize()
                self.compute.synchronize()
                for ln in self.lanes:
                    ln.plan = _Plan(m.image_size[0], m.num_joints, m.num_classes, self.batch, params)
            self._params = params

    def _forward_decode(self, ln, dt, st):
        """forward + keypoint decode of lane `ln` on the compute stream `st`."""
        lib = _lib.load()
        _lib.check(lib.hgr_forward_keypoints(ln.plan.handle, ln.d_x.data_ptr(), dt, self.batch, ln.d_logits.data_ptr(),
                                             ln.d_heat.data_ptr() if ln.d_heat is not None else None,
                                             ln.d_preds.data_ptr(), ln.d_maxvals.data_ptr(), _lib.F32, st),
                   "hgr_forward_keypoints")

    def _copy_back(self, ln):
        """results device -> pinned host on the lane's copy stream, after the compute stream is done with them."""
        with torch.cuda.stream(ln.stream):
            ln.stream.wait_event(ln.compute_done)
            ln.h_logits.copy_(ln.d_logits, non_blocking=True)
            ln.h_preds.copy_(ln.d_preds, non_blocking=True)
            ln.h_maxvals.copy_(ln.d_maxvals, non_blocking=True)
            ln.done.record(ln.stream)

    def submit(self, crops_u8: torch.Tensor, after: torch.cuda.Event | None = None) -> int:
        """Queue one batch of (B, S, S, 3) uint8 host crops; returns the lane to collect from."""
        lib = _lib.load()
        self._refresh_plans()
        i = self._next
        self._next = (self._next + 1) % len(self.lanes)
        ln = self.lanes[i]
        if ln.busy:
            raise RuntimeError("lane still holds an uncollected batch; call collect() first")
        if tuple(crops_u8.shape) != tuple(ln.h_crops.shape) or crops_u8.dtype != torch.uint8 or crops_u8.is_cuda:
            raise ValueError(f"expected host uint8 crops of shape {tuple(ln.h_crops.shape)}")
        dt = _lib.F32 if self.dtype == torch.float32 else _lib.BF16
        m = self.model
        s = m.image_size[0]
        with torch.cuda.device(self.device):
            with torch.cuda.stream(ln.stream):
                if after is not None:
                    ln.stream.wait_event(after)
                src = crops_u8 if crops_u8.is_pinned() else ln.h_crops.copy_(crops_u8)
                ln.d_crops.copy_(src, non_blocking=True)
                ln.h2d_done.record(ln.stream)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(ln.h2d_done)
                st = self.compute.cuda_stream
                _lib.check(lib.hgr_crop_normalize(ln.d_crops.data_ptr(), ln.d_x.data_ptr(), dt, self.batch, s, s, st),
                           "hgr_crop_normalize")
                self._forward_decode(ln, dt, st)
                ln.compute_done.record(self.compute)
            self._copy_back(ln)
        ln.busy = True
        return i

    def submit_frames(self, frames_u8: torch.Tensor, boxes, frame_index=None) -> int:
        """The whole classifier stage of detect.py:119-155 for `batch` detector boxes: frames (F, Hf, Wf, 3) uint8
        in HOST memory are copied to the device, every box is warped to the crop size with cv2.warpAffine's
        arithmetic and normalised in ONE kernel (crop_warp_normalize), then forward + keypoint decode as in submit()."""
        from .ops import box_to_affine, invert_affine
        import numpy as np
        lib = _lib.load()
        self._refresh_plans()
        i = self._next
        self._next = (self._next + 1) % len(self.lanes)
        ln = self.lanes[i]
        if ln.busy:
            raise RuntimeError("lane still holds an uncollected batch; call collect() first")
        if len(boxes) != self.batch:
            raise ValueError(f"expected {self.batch} boxes (the plan's batch), got {len(boxes)}")
        if frames_u8.dtype != torch.uint8 or frames_u8.is_cuda or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
            raise ValueError("expected host uint8 frames (F, Hf, Wf, 3)")
        m = self.model
        s = m.image_size[0]
        idx = np.zeros(self.batch, dtype=np.int32) if frame_index is None else np.asarray(frame_index, dtype=np.int32)
        # the kernel indexes the frames buffer with these: a bad index would read past its end (ops.crop_warp_normalize
        # checks the same)
        if idx.shape != (self.batch,) or (idx < 0).any() or (idx >= frames_u8.shape[0]).any():
            raise ValueError(f"frame_index must hold {self.batch} indices in [0, {frames_u8.shape[0]})")
        inv = np.stack([invert_affine(box_to_affine(b, s)) for b in boxes])
        dt = _lib.F32 if self.dtype == torch.float32 else _lib.BF16
        with torch.cuda.device(self.device):
            with torch.cuda.stream(ln.stream):
                ln.d_frames = frames_u8.contiguous().to(self.device, non_blocking=True)
                ln.d_inv = torch.from_numpy(np.ascontiguousarray(inv)).to(self.device, non_blocking=True)
                ln.d_idx = torch.from_numpy(idx).to(self.device, non_blocking=True)
                ln.h2d_done.record(ln.stream)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(ln.h2d_done)
                st = self.compute.cuda_stream
                _lib.check(lib.hgr_crop_warp_normalize(ln.d_frames.data_ptr(), frames_u8.shape[0], frames_u8.shape[1],
                                                       frames_u8.shape[2], ln.d_idx.data_ptr(), ln.d_inv.data_ptr(),
                                                       self.batch, s, ln.d_x.data_ptr(), dt, st),
                           "hgr_crop_warp_normalize")
                self._forward_decode(ln, dt, st)
                ln.compute_done.record(self.compute)
            self._copy_back(ln)
        ln.busy = True
        return i

    def infer_frames(self, frames_u8: torch.Tensor, boxes, frame_index=None):
        a, b, c = self.collect(self.submit_frames(frames_u8, boxes, frame_index))
        return a.clone(), b.clone(), c.clone()

    def collect(self, lane: int):
        """Wait for a submitted batch; returns host tensors (logits, preds, maxvals) valid until the lane is reused."""
        ln = self.lanes[lane]
        if not ln.busy:
            raise RuntimeError("nothing submitted on this lane")
        ln.done.synchronize()
        ln.busy = False
        return ln.h_logits, ln.h_preds, ln.h_maxvals

    def infer(self, crops_u8: torch.Tensor):
        """One batch, synchronously: (logits (B, C), keypoints (B, J, 2), confidences (B, J, 1)) on the host."""
        a, b, c = self.collect(self.submit(crops_u8))
        return a.clone(), b.clone(), c.clone()


class ShardedHandPipeline:
    """One process, several GPUs: BASELINE.json configs[2] ("batch 8192 sharded across 1/2/4/8 B200, no collective")
    as a host-facing call.  A host batch of crops is cut into contiguous shards with `sharding.shard_range`, every
    device runs its own replica of the model through its own HandPipeline (copy stream + compute stream + plan per
    device), all shards are submitted before the first one is collected, and the results come back in batch order.
    There is no cross-device traffic: every crop is independent (reference detect.py:119-155 classifies one crop at a
    time).  The host buffers are pinned per device pipeline; the process is expected to be bound to the NUMA node of
    its GPUs by the launcher (numactl) - PyTorch offers no per-allocation NUMA placement."""

    def __init__(self, model: MultiTaskNet, devices, total_batch: int, compute_dtype=torch.bfloat16):
        import copy
        from .sharding import shard_range
        devices = [torch.device(d) for d in devices]
        if not devices or any(d.type != "cuda" for d in devices):
            raise RuntimeError("ShardedHandPipeline needs CUDA devices; there is no CPU path")
        if model.training:
            raise RuntimeError("ShardedHandPipeline needs model.eval()")
        self.total_batch = int(total_batch)
        self.ranges = [shard_range(self.total_batch, r, len(devices)) for r in range(len(devices))]
        self.pipes = []
        home = next(model.parameters()).device
        for d, (lo, hi) in zip(devices, self.ranges):
            if hi == lo:
                self.pipes.append(None)
                continue
            replica = model if d == home else copy.deepcopy(model).to(d).eval()
            replica.return_attention = model.return_attention
            self.pipes.append(HandPipeline(replica, hi - lo, compute_dtype, lanes=1))

    def infer(self, crops_u8: torch.Tensor):
        """(total_batch, S, S, 3) uint8 host crops -> (logits, keypoints, confidences) on the host, in batch order."""
        if crops_u8.shape[0] != self.total_batch or crops_u8.is_cuda or crops_u8.dtype != torch.uint8:
            raise ValueError(f"expected {self.total_batch} host uint8 crops")
        lanes = [p.submit(crops_u8[lo:hi]) if p is not None else None for p, (lo, hi) in zip(self.pipes, self.ranges)]
        outs = [p.collect(ln) for p, ln in zip(self.pipes, lanes) if p is not None]
        return tuple(torch.cat([o[k] for o in outs]) for k in range(3))
