"""Drop-in replacement for the reference's `model.multitasknet.MultiTaskNet`.

Same constructor `MultiTaskNet(num_joints, num_classes, image_size)`, same
module tree and therefore the same 180 `state_dict` keys (reference
model/multitasknet.py:8-29, model/gelan.py:18-176, model/transformer.py:45-127),
same `forward(x) -> (cls_out, hmap_out, attnmap)` contract, in `.eval()` (BatchNorm folded into the
convolutions) and in `.train()` (batch statistics, running-stat updates, autograd).  The sub-modules
are parameter containers only: all arithmetic happens in the hand-written
sm_100a kernels of libhgr_b200.so, reached through the C ABI in
include/hgr_b200.h.  There is no CPU path and no PyTorch-operator path:
a CPU tensor, a missing library or an unsupported mode raises.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib, packing

_HEADS, _HEAD_DIM, _DIM, _DEPTH, _MLP = 8, 32, 256, 4, 256


class _Container(nn.Module):
    """A node of the module tree that only holds parameters."""

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container of the B200 MultiTaskNet; "
            "call MultiTaskNet.forward, which runs the fused CUDA path")


class Conv(_Container):
    """conv (no bias) + BatchNorm2d [+ SiLU]: holds `conv.weight` and the `bn.*` entries."""

    def __init__(self, c1, c2, k=1, s=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.has_act = bool(act)


class ResBasicBlock(_Container):
    """SiLU(x + BN(conv3(SiLU(BN(conv3(x))))))  (c1 == c2 everywhere in 'small')."""

    def __init__(self, c1, c2):
        super().__init__()
        if c1 != c2:
            raise ValueError("the B200 path implements the identity-shortcut block only (c1 == c2)")
        self.cv1 = Conv(c1, c2, 3, 1)
        self.cv2 = Conv(c2, c2, 3, 1, act=False)


class ResBottleneck(_Container):
    """The reference's 1x1 -> 3x3 -> 1x1 bottleneck (model/gelan.py:90-121): same constructor and the same
    `cv1` / `cv2` / `cv3` parameter tree, so that code which builds or loads one keeps working.  `gelan_spec`
    (gelan.py:148-151) never instantiates it and `MultiTaskNet` hard-codes 'small' (multitasknet.py:12), so no
    kernel path exists for it: like every container here it holds parameters and raises when called."""

    def __init__(self, c1, c2, shortcut=True, e=0.5):
        super().__init__()
        c_ = int(c2 * e)  # hidden channels
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_, c_, 3, 1)
        self.cv3 = Conv(c_, c2, 1, 1, act=False)
        self.add = bool(shortcut and c1 == c2)
        self.downsample = None  # the reference only creates one when add and c1 != c2, which cannot both hold


class GELANBlock(_Container):
    def __init__(self, c_in, c_out, c_hid1, c_hid2, nblocks=1):
        super().__init__()
        if nblocks != 1:
            raise ValueError("only GELANNet('small') (one block per stack) is implemented")
        self.cv1 = Conv(c_in, c_hid1, 1, 1)
        self.cv2 = nn.Sequential(ResBasicBlock(c_hid1 // 2, c_hid2))
        self.cv3 = nn.Sequential(ResBasicBlock(c_hid2, c_hid2))
        self.cv4 = Conv(c_hid1 + 2 * c_hid2, c_out, 1, 1)


class GELANNet(_Container):
    def __init__(self, gelan_type="small"):
        super().__init__()
        if gelan_type != "small":
            raise ValueError("MultiTaskNet hard-codes GELANNet('small') (reference multitasknet.py:12)")
        self.conv1 = Conv(3, 64, 3, 2)
        self.conv2 = Conv(64, 128, 3, 2)
        self.cspelan1 = GELANBlock(128, 128, 128, 64)
        self.down1 = Conv(128, 256, 3, 2)
        self.cspelan2 = GELANBlock(256, 256, 256, 128)
        self.down2 = Conv(256, 512, 3, 2)
        self.cspelan3 = GELANBlock(512, 512, 512, 256)


class Attention(_Container):
    def __init__(self, dim, heads, head_dim):
        super().__init__()
        self.heads = heads
        self.scale = head_dim ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.to_qkv = nn.Linear(dim, heads * head_dim * 3, bias=False)
        self.to_out = nn.Linear(heads * head_dim, dim, bias=False)


class FeedForward(_Container):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        # indices 0, 1 and 4 carry parameters, exactly like the reference's Sequential
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Transformer(_Container):
    def __init__(self, dim, depth, heads, head_dim, mlp_dim):
        super().__init__()
        self.layers = nn.ModuleList(
            nn.ModuleList([Attention(dim, heads, head_dim), FeedForward(dim, mlp_dim)]) for _ in range(depth))


class ViT(_Container):
    def __init__(self, num_classes, num_joints, feature_size, dim=_DIM, depth=_DEPTH, heads=_HEADS,
                 head_dim=_HEAD_DIM, mlp_dim=_MLP):
        super().__init__()
        # plain attribute, not a buffer: it is not part of the state_dict (reference transformer.py:103-107)
        self.pos_embedding = packing.sincos_table(feature_size[0], feature_size[1], dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.transformer = Transformer(dim, depth, heads, head_dim, mlp_dim)
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))
        self.simple_decoder = nn.Sequential(nn.ReLU(inplace=True), nn.Conv2d(dim, num_joints, 1))


class _Plan:
    """A bound hgr_plan plus the torch tensors that own its memory."""

    def __init__(self, image_size, num_joints, num_classes, batch, params: torch.Tensor):
        lib = _lib.load()
        nbytes = lib.hgr_workspace_bytes(image_size, batch)
        if nbytes == 0:
            _lib.check(-1, "hgr_workspace_bytes")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=params.device)
        self.params = params
        self.batch = batch
        handle = C.c_void_p()
        _lib.check(lib.hgr_plan_create(C.byref(handle), image_size, num_joints, num_classes, batch,
                                       params.data_ptr(), self.workspace.data_ptr(), nbytes), "hgr_plan_create")
        self.handle = handle

    def buffer(self, name: str) -> torch.Tensor:
        """bf16 NHWC view of a named intermediate (for per-stage parity tests)."""
        ptr = C.c_void_p()
        dims = (C.c_int64 * 4)()
        _lib.check(_lib.load().hgr_plan_buffer(self.handle, name.encode(), C.byref(ptr), dims), "hgr_plan_buffer")
        off = ptr.value - self.workspace.data_ptr()
        n = dims[0] * dims[1] * dims[2] * dims[3]
        return self.workspace[off: off + 2 * n].view(torch.bfloat16).view(*dims)

    def launches(self) -> int:
        return _lib.load().hgr_plan_launches(self.handle, 0)

    def launch_table(self):
        """[(layer name, kind, algorithmic flops, algorithmic bytes)] of one forward pass."""
        lib = _lib.load()
        out = []
        for i in range(self.launches()):
            name = C.c_char_p()
            kind = C.c_int()
            fl, by = C.c_double(), C.c_double()
            _lib.check(lib.hgr_plan_launch_info(self.handle, i, C.byref(name), C.byref(kind), C.byref(fl),
                                                C.byref(by)), "hgr_plan_launch_info")
            out.append((name.value.decode(), kind.value, fl.value, by.value))
        return out

    def profile(self, x, logits, heat, attn=None):
        """Per-launch milliseconds of one forward (CUDA events between launches)."""
        n = self.launches()
        ms = (C.c_float * n)()
        dt = _lib.F32 if x.dtype == torch.float32 else _lib.BF16
        odt = _lib.F32 if logits.dtype == torch.float32 else _lib.BF16
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rc = _lib.load().hgr_forward_profile(self.handle, x.data_ptr(), dt, x.shape[0], logits.data_ptr(),
                                             heat.data_ptr(), attn.data_ptr() if attn is not None else None, odt,
                                             stream, ms, n)
        if rc != n:
            _lib.check(rc if rc < 0 else -1, "hgr_forward_profile")
        return list(ms)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().hgr_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class MultiTaskNet(nn.Module):
    """GELAN backbone -> 1x1 proj -> 4-layer ViT -> (gesture logits, pose heatmaps, last attention map).

    `return_attention=False` (attribute, not a constructor argument so the
    constructor stays the reference's) skips materialising the (B, 8, T, T)
    probabilities and returns None in their place; the reference's callers
    tolerate that (libs/vis.py:204, export.py:44).
    """

    def __init__(self, num_joints, num_classes, image_size):
        super().__init__()
        if image_size[0] != image_size[1]:
            raise ValueError("square inputs only (the reference enforces it at train.py:191-192)")
        self.num_joints, self.num_classes = int(num_joints), int(num_classes)
        self.image_size = [int(image_size[0]), int(image_size[1])]
        self.encoder = GELANNet("small")
        self.proj = nn.Conv2d(512, 256, 1, bias=False)
        self.decoder = ViT(num_classes=num_classes, num_joints=num_joints,
                           feature_size=[image_size[0] // 16, image_size[1] // 16])
        self.return_attention = True
        self._packed = None        # (signature, uint8 tensor)
        self._plans = {}           # (device index, batch) -> _Plan

    # ---- weight packing ---------------------------------------------------
    def _signature(self):
        """(storage address, version counter) of every parameter and buffer: changes when a tensor is written in
        place (load_state_dict, optimiser step), re-homed (.to()) or replaced (load_state_dict(assign=True)).
        The module tree is fixed, so the per-module `_parameters` / `_buffers` dicts are collected once and only
        their current values are read here (~50 us instead of ~500 us for a state_dict walk per forward)."""
        dicts = self.__dict__.get("_sig_dicts")
        if dicts is None:
            dicts = [d for m in self.modules() for d in (m._parameters, m._buffers) if d]
            self.__dict__["_sig_dicts"] = dicts
        return tuple((t.data_ptr(), t._version) for d in dicts for t in d.values() if t is not None)

    def _packed_params(self, device) -> torch.Tensor:
        sig = (device, self._signature())
        if self._packed is None or self._packed[0] != sig:
            block = packing.pack(self.state_dict(), self.image_size[0], self.num_joints, self.num_classes, device)
            self._packed = (sig, block)
            self._plans.clear()
        return self._packed[1]

    def plan_for(self, batch: int, device) -> _Plan:
        params = self._packed_params(device)
        key = (device.index, batch)
        if key not in self._plans:
            self._plans[key] = _Plan(self.image_size[0], self.num_joints, self.num_classes, batch, params)
        return self._plans[key]

    # ---- forward ------------------------------------------------------------
    def forward(self, x):
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise TypeError("expected a (B, 3, S, S) tensor")
        if not x.is_cuda:
            raise RuntimeError("the B200 MultiTaskNet has no CPU path: move the input (and the module) to a CUDA device")
        s = self.image_size[0]
        if tuple(x.shape[1:]) != (3, s, s):
            raise ValueError(f"input {tuple(x.shape)} does not match image_size {self.image_size} "
                             "(the reference fails on the position-embedding broadcast)")
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("input must be float32 or bfloat16")
        if next(self.parameters()).device != x.device:
            raise RuntimeError("module parameters and input are on different devices")
        b = x.shape[0]
        f = s // 16
        t = f * f + 1
        x = x.contiguous()
        if self.training:
            # batch-statistics BatchNorm + autograd (training.py): the parameters become views of one flat block
            from . import training
            return training.forward_autograd(self, x)
        dt = _lib.F32 if x.dtype == torch.float32 else _lib.BF16
        with torch.cuda.device(x.device):
            plan = self.plan_for(b, x.device)
            cls_out = torch.empty(b, self.num_classes, dtype=x.dtype, device=x.device)
            hmap_out = torch.empty(b, self.num_joints, s // 4, s // 4, dtype=x.dtype, device=x.device)
            attn = torch.empty(b, _HEADS, t, t, dtype=x.dtype, device=x.device) if self.return_attention else None
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(_lib.load().hgr_forward(plan.handle, x.data_ptr(), dt, b, cls_out.data_ptr(),
                                               hmap_out.data_ptr(), attn.data_ptr() if attn is not None else None,
                                               dt, stream), "hgr_forward")
        return cls_out, hmap_out, attn
