#!/usr/bin/env python
"""Timeline of the fused ViT layer-tail kernel (csrc/vit_block.cu): clock64 marks of CTA 0 for a few tiles at the
batch-1024 shape (148 480 token rows), printed as cycle deltas between the hand-over points of the chain."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "hand-gesture-recognition_b200"):
    sys.path.insert(0, str(p))
import torch
from hgr_b200 import _lib

rows, nt = 1024 * 145, 8
dev = torch.device("cuda", 0)
lib = _lib.load()
g = torch.Generator().manual_seed(0)
bf = lambda *s: (torch.randn(*s, generator=g) / (16 if len(s) == 2 and s[0] == 256 else 1)).to(dev, torch.bfloat16)
a, x0, wo, w1, w2 = bf(rows, 256), bf(rows, 256), bf(256, 256), bf(256, 256), bf(256, 256)
c1, d1, b2 = (torch.randn(256, generator=g).to(dev) for _ in range(3))
x2 = torch.empty_like(x0)
stats = torch.empty(rows, 2, device=dev)
trace = torch.zeros(nt, 16, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.hgr_vit_block_trace(a.data_ptr(), x0.data_ptr(), rows, wo.data_ptr(), w1.data_ptr(), c1.data_ptr(),
                                       d1.data_ptr(), w2.data_ptr(), b2.data_ptr(), x2.data_ptr(), stats.data_ptr(),
                                       trace.data_ptr(), nt, st), "hgr_vit_block_trace")
torch.cuda.synchronize()
t = trace.cpu()
base = int(t[0, 8])
names = {6: "epi at top", 11: "w0", 12: "w1", 13: "w2", 14: "w3", 8: "G0 operand ready", 0: "G0 done", 7: "x0 landed", 1: "E0 done", 9: "G1 start", 2: "G1 done", 3: "E1 done", 10: "G2 start",
         4: "G2 done", 5: "E2 done"}
order = [6, 8, 14, 0, 7, 1, 9, 2, 3, 10, 4, 5]
for i in range(nt):
    prev = None
    parts = []
    for e in order:
        v = int(t[i, e]) - base
        parts.append(f"{names[e]} {v}" + (f" (+{v - prev})" if prev is not None else ""))
        prev = v
    print(f"tile {i}: " + " | ".join(parts))
