#!/usr/bin/env python
"""Runs ONE operator of the C ABI a few times at the headline batch (development tool: the command ncu wraps when a
single kernel is profiled, and a quick CUDA-event timer).  Usage: tools/run_op.py <attention_tc|attention|pose_head|conv1|stem_fused|stem_two|gelan_tail|gelan_tail_two> [B] [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hand-gesture-recognition_b200"))
import torch
from hgr_b200 import _lib

op = sys.argv[1] if len(sys.argv) > 1 else "attention_tc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lib = _lib.load()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
T, F, J = 145, 12, 21
if op in ("attention_tc", "attention"):
    qkv = (torch.randn(B, T, 768, generator=g, device=dev) * 1.5).bfloat16()
    out = torch.empty(B, T, 256, dtype=torch.bfloat16, device=dev)
    if op == "attention_tc":
        run = lambda: _lib.check(lib.hgr_attention_tc(qkv.data_ptr(), out.data_ptr(), B, T, st), op)
    else:
        run = lambda: _lib.check(lib.hgr_attention(qkv.data_ptr(), out.data_ptr(), None, _lib.F32, B, T, st), op)
elif op == "pose_head":
    tok = torch.randn(B, T, 256, generator=g, device=dev).bfloat16()
    w = (torch.randn(J, 256, generator=g, device=dev) * 0.06).bfloat16()
    bias = torch.randn(J, generator=g, device=dev)
    heat = torch.empty(B, J, 4 * F, 4 * F, dtype=torch.bfloat16, device=dev)
    run = lambda: _lib.check(lib.hgr_pose_head(tok.data_ptr(), w.data_ptr(), bias.data_ptr(), heat.data_ptr(), _lib.BF16,
                                               B, F, J, st), op)
elif op == "conv1":
    S = 192
    x = torch.randn(B, 3, S, S, generator=g, device=dev).bfloat16()
    wk = torch.zeros(64, 32, device=dev)
    wk[:, :27] = torch.randn(64, 27, generator=g, device=dev) * 0.27
    wk = wk.bfloat16()
    sh = torch.randn(64, generator=g, device=dev) * 0.3
    out = torch.empty(B, S // 2, S // 2, 64, dtype=torch.bfloat16, device=dev)
    run = lambda: _lib.check(lib.hgr_conv1(x.data_ptr(), _lib.BF16, B, S, wk.data_ptr(), sh.data_ptr(), out.data_ptr(), st), op)
elif op in ("stem_fused", "stem_two"):
    S = 192
    x = torch.randn(B, 3, S, S, generator=g, device=dev).bfloat16()
    wk = torch.zeros(64, 32, device=dev)
    wk[:, :27] = torch.randn(64, 27, generator=g, device=dev) * 0.27
    wk = wk.bfloat16()
    sh = torch.randn(64, generator=g, device=dev) * 0.3
    w1 = (torch.randn(128, 9, 64, generator=g, device=dev) * 0.06).bfloat16()
    w2 = (torch.randn(128, 128, generator=g, device=dev) * 0.12).bfloat16()
    s1, s2 = torch.rand(128, generator=g, device=dev) + 0.5, torch.rand(128, generator=g, device=dev) + 0.5
    t1, t2 = torch.randn(128, generator=g, device=dev) * 0.3, torch.randn(128, generator=g, device=dev) * 0.3
    a1 = torch.empty(B, S // 2, S // 2, 64, dtype=torch.bfloat16, device=dev)
    out = torch.empty(B, S // 4, S // 4, 256, dtype=torch.bfloat16, device=dev)
    if op == "stem_fused":
        run = lambda: _lib.check(lib.hgr_stem_fused(x.data_ptr(), B, S, wk.data_ptr(), sh.data_ptr(), w1.data_ptr(),
                                                    s1.data_ptr(), t1.data_ptr(), w2.data_ptr(), s2.data_ptr(),
                                                    t2.data_ptr(), out.data_ptr(), 256, 0, st), op)
    else:
        def run():
            _lib.check(lib.hgr_conv1(x.data_ptr(), _lib.BF16, B, S, wk.data_ptr(), sh.data_ptr(), a1.data_ptr(), st), op)
            _lib.check(lib.hgr_conv_chain(a1.data_ptr(), B, S // 2, S // 2, w1.data_ptr(), s1.data_ptr(), t1.data_ptr(),
                                          w2.data_ptr(), s2.data_ptr(), t2.data_ptr(), out.data_ptr(), 256, 0, st), op)
elif op in ("gelan_tail", "gelan_tail_two"):
    H = 48
    t = torch.randn(B, H, H, 64, generator=g, device=dev).bfloat16()
    gb = torch.randn(B, H, H, 256, generator=g, device=dev).bfloat16()
    wh = (torch.randn(64, 9, 64, generator=g, device=dev) * 0.06).bfloat16()
    w4 = (torch.randn(128, 256, generator=g, device=dev) * 0.09).bfloat16()
    sh_, s4 = torch.rand(64, generator=g, device=dev) + 0.5, torch.rand(128, generator=g, device=dev) + 0.5
    th_, t4 = torch.randn(64, generator=g, device=dev) * 0.3, torch.randn(128, generator=g, device=dev) * 0.3
    out = torch.empty(B, H, H, 128, dtype=torch.bfloat16, device=dev)
    if op == "gelan_tail":
        run = lambda: _lib.check(lib.hgr_gelan_tail(t.data_ptr(), gb.data_ptr(), B, H, H, wh.data_ptr(), sh_.data_ptr(),
                                                    th_.data_ptr(), w4.data_ptr(), s4.data_ptr(), t4.data_ptr(),
                                                    out.data_ptr(), st), op)
    else:
        def run():
            _lib.check(lib.hgr_conv_bn_act(t.data_ptr(), B, H, H, 64, 0, 64, wh.data_ptr(), sh_.data_ptr(), th_.data_ptr(),
                                           3, 1, 1, gb.data_ptr(), 256, 128, gb.data_ptr(), 256, 192, 64, st), op)
            _lib.check(lib.hgr_conv_bn_act(gb.data_ptr(), B, H, H, 256, 0, 256, w4.data_ptr(), s4.data_ptr(), t4.data_ptr(),
                                           1, 1, 1, None, 0, 0, out.data_ptr(), 128, 0, 128, st), op)
else:
    raise SystemExit(f"unknown op {op}")
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
print(f"{op} B={B}: {e0.elapsed_time(e1) / reps:.4f} ms per launch")
