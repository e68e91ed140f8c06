#!/usr/bin/env python
"""Cycle timeline of CTA 0 of the tcgen05 attention kernel (development tool).
Usage: tools/attention_trace.py [B] [items]   -> prints, per item and unit position r, when each phase happened
relative to the first mark, so that the overlap of the three units in flight can be read off."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hand-gesture-recognition_b200"))
import torch
from hgr_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
items = int(sys.argv[2]) if len(sys.argv) > 2 else 12
lib = _lib.load()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
T = 145
qkv = (torch.randn(B, T, 768, generator=g, device=dev) * 1.5).bfloat16()
out = torch.empty(B, T, 256, dtype=torch.bfloat16, device=dev)
warps = C.c_int()
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.hgr_attention_tc_trace(None, None, B, T, None, 0, C.byref(warps), st), "size query")
W = warps.value
for rep in range(3):
    trace = torch.zeros(items, W, 8, dtype=torch.int64, device=dev)
    _lib.check(lib.hgr_attention_tc_trace(qkv.data_ptr(), out.data_ptr(), B, T, trace.data_ptr(), items,
                                          C.byref(warps), st), "trace")
    torch.cuda.synchronize()
tr = trace.cpu()
t0 = int(tr[tr > 0].min())
rel = lambda v: int(v) - t0 if v > 0 else -1
print(f"warps {W}; times in cycles since the first mark")
for i in range(items):
    m = tr[i, 1]
    print(f"item {i}: MMA scores issued r0/r1/r2 {rel(m[0])}/{rel(m[1])}/{rel(m[2])}  PV issued {rel(m[3])}/{rel(m[4])}/{rel(m[5])}")
    for u in range(3 * i, 3 * i + 3):
        r, grp = u % 3, u % 2
        ws = range(8 + 4 * grp, 12 + 4 * grp)
        tb = 4 if r == 2 else 0
        def col(k, f):
            vals = [int(tr[i, w, tb + k]) for w in ws if tr[i, w, tb + k] > 0]
            return f(vals) - t0 if vals else -1
        dr = [int(tr[i, w, 2 * r]) for w in range(4, 8) if tr[i, w, 2 * r] > 0]
        dd = [int(tr[i, w, 2 * r + 1]) for w in range(4, 8) if tr[i, w, 2 * r + 1] > 0]
        print(f"   unit r{r} (group {grp}): scores {col(0, min)}  max_done {col(1, min)}..{col(1, max)}  P_done {col(2, min)}..{col(2, max)}"
              f"  | drain O_ready {min(dr) - t0 if dr else -1} stored {max(dd) - t0 if dd else -1}")
