"""ctypes binding of libhgr_b200.so (the C ABI declared in include/hgr_b200.h).

The library is built in-tree by build.py; importing this module never falls
back to another implementation: if the shared object is missing or does not
load, every call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2

_LIB_PATH = Path(__file__).resolve().parent / "libhgr_b200.so"
_lib = None

_vp, _i, _ll, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t
_fp = C.c_void_p  # float* passed as raw addresses

# name -> (restype, argtypes); mirrors include/hgr_b200.h one to one
SIGNATURES = {
    "hgr_version": (_i, []),
    "hgr_last_error": (C.c_char_p, []),
    "hgr_param_count": (_i, [_i, _i, _i]),
    "hgr_param_info": (_i, [_i, _i, _i, _i, C.POINTER(C.c_char_p), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i),
                            C.POINTER(C.c_int64)]),
    "hgr_param_bytes": (_sz, [_i, _i, _i]),
    "hgr_workspace_bytes": (_sz, [_i, _i]),
    "hgr_plan_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _vp, _vp, _sz]),
    "hgr_plan_destroy": (None, [_vp]),
    "hgr_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "hgr_keypoints_fused": (_i, [_i, _i]),
    "hgr_forward_keypoints": (_i, [_vp, _vp, _i, _i, _vp, _vp, _fp, _fp, _i, _vp]),
    "hgr_forward_host": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "hgr_plan_buffer": (_i, [_vp, C.c_char_p, C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "hgr_plan_launches": (_i, [_vp, _i]),
    "hgr_plan_launch_info": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_i), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double)]),
    "hgr_forward_profile": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, C.POINTER(C.c_float), _i]),
    "hgr_conv_bn_act": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _fp, _fp, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _i,
                             _vp]),
    "hgr_conv_chain": (_i, [_vp, _i, _i, _i, _vp, _fp, _fp, _vp, _fp, _fp, _vp, _i, _i, _vp]),
    "hgr_stem_fused": (_i, [_vp, _i, _i, _vp, _fp, _vp, _fp, _fp, _vp, _fp, _fp, _vp, _i, _i, _vp]),
    "hgr_gelan_tail": (_i, [_vp, _vp, _i, _i, _i, _vp, _fp, _fp, _vp, _fp, _fp, _vp, _vp]),
    "hgr_linear": (_i, [_vp, _ll, _i, _vp, _fp, _fp, _i, _vp, _vp, _i, _fp, _fp, _vp]),
    "hgr_vit_block": (_i, [_vp, _vp, _ll, _vp, _vp, _fp, _fp, _vp, _fp, _vp, _fp, _vp]),
    "hgr_vit_block_trace": (_i, [_vp, _vp, _ll, _vp, _vp, _fp, _fp, _vp, _fp, _vp, _fp, _vp, _i, _vp]),
    "hgr_conv1": (_i, [_vp, _i, _i, _i, _vp, _fp, _vp, _vp]),
    "hgr_layernorm": (_i, [_vp, _vp, _fp, _fp, _ll, _vp]),
    "hgr_attention": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "hgr_attention_tc": (_i, [_vp, _vp, _i, _i, _vp]),
    "hgr_attention_tc_trace": (_i, [_vp, _vp, _i, _i, _vp, _i, C.POINTER(_i), _vp]),
    "hgr_cls_head": (_i, [_vp, _fp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _vp]),
    "hgr_pose_head": (_i, [_vp, _vp, _fp, _vp, _i, _i, _i, _i, _vp]),
    "hgr_pose_head_decode": (_i, [_vp, _vp, _fp, _vp, _i, _fp, _fp, _i, _i, _i, _vp]),
    "hgr_get_max_preds": (_i, [_vp, _i, _i, _i, _i, _i, _fp, _fp, _vp]),
    "hgr_crop_normalize": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "hgr_pose_accuracy": (_i, [_vp, _vp, _i, _i, _i, _i, C.c_double, _vp, _vp, _vp, _vp]),
    "hgr_crop_warp_normalize": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _vp]),
    # ---- training step ----
    "hgr_train_param_count": (_i, [_i, _i]),
    "hgr_train_param_info": (_i, [_i, _i, _i, C.POINTER(C.c_char_p), C.POINTER(_sz), C.POINTER(_sz)]),
    "hgr_train_param_floats": (_sz, [_i, _i]),
    "hgr_train_bnstat_count": (_i, []),
    "hgr_train_bnstat_info": (_i, [_i, C.POINTER(C.c_char_p), C.POINTER(_sz), C.POINTER(_sz)]),
    "hgr_train_bnstat_floats": (_sz, []),
    "hgr_train_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "hgr_train_plan_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz]),
    "hgr_train_plan_destroy": (None, [_vp]),
    "hgr_train_buffer": (_i, [_vp, C.c_char_p, C.POINTER(_vp), C.POINTER(_sz), C.POINTER(C.c_int64)]),
    "hgr_train_forward": (_i, [_vp, _vp, _i, _vp, _vp, C.c_float, _vp]),
    "hgr_train_backward": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "hgr_train_backward_part": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp]),
    "hgr_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, C.c_float, _vp, _vp, _vp, _vp, _vp]),
    "hgr_adamw_step": (_i, [_vp, _vp, _vp, _vp, _ll, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _i,
                            C.c_float, _vp]),
    "hgr_wgrad": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "hgr_wgrad_partial_floats": (_sz, [_i, _i, _i, _ll]),
    "hgr_attention_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "hgr_dgrad_s2": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp]),
}


class HgrError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise HgrError(
            f"{_LIB_PATH} is missing: build it with `python __graft_entry__.py` or "
            "`python hand-gesture-recognition_b200/hgr_b200/build.py` (needs nvcc). "
            "There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hgr_last_error()
        raise HgrError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def param_layout(image_size: int, num_joints: int, num_classes: int):
    """[(name, offset, nbytes, dtype_code, dims)] of the packed parameter block."""
    lib = load()
    n = lib.hgr_param_count(image_size, num_joints, num_classes)
    if n < 0:
        check(n, "hgr_param_count")
    out = []
    for i in range(n):
        name = C.c_char_p()
        off, nb = _sz(), _sz()
        dt = _i()
        dims = (C.c_int64 * 3)()
        check(lib.hgr_param_info(image_size, num_joints, num_classes, i, C.byref(name), C.byref(off), C.byref(nb),
                                 C.byref(dt), dims), "hgr_param_info")
        out.append((name.value.decode(), off.value, nb.value, dt.value, tuple(dims)))
    return out


def train_param_layout(num_joints: int, num_classes: int):
    """[(state_dict key, offset in floats, numel)] of the flat training parameter / gradient block."""
    lib = load()
    out = []
    for i in range(lib.hgr_train_param_count(num_joints, num_classes)):
        name = C.c_char_p()
        off, n = _sz(), _sz()
        check(lib.hgr_train_param_info(num_joints, num_classes, i, C.byref(name), C.byref(off), C.byref(n)),
              "hgr_train_param_info")
        out.append((name.value.decode(), off.value, n.value))
    return out


def train_bnstat_layout():
    """[(state_dict key, offset in floats, numel)] of the flat BatchNorm running-statistics block."""
    lib = load()
    out = []
    for i in range(lib.hgr_train_bnstat_count()):
        name = C.c_char_p()
        off, n = _sz(), _sz()
        check(lib.hgr_train_bnstat_info(i, C.byref(name), C.byref(off), C.byref(n)), "hgr_train_bnstat_info")
        out.append((name.value.decode(), off.value, n.value))
    return out
