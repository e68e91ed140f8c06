#!/usr/bin/env python
"""How the implicit-GEMM kernel's time scales with K and N for the ViT's token count (development probe behind
DESIGN.md section 9, row 1): rows = 1024 x 145, y = x W^T.  Measured on B200: 0.047 ms at K = 64 (the store / epilogue
floor, 5.2 TB/s), + ~0.145 us per unit of K (0.0725 ms at K = 256, 0.117 at 512, 0.184 at 1024): the two parts ADD."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hand-gesture-recognition_b200"))
import torch
from hgr_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda"); st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev).manual_seed(0)
rows = 1024 * 145
def run(K, N, stats, reps=20):
    x = torch.randn(rows, K, generator=g, device=dev).bfloat16()
    w = (torch.randn(N, K, generator=g, device=dev) / 16).bfloat16()
    c = torch.randn(N, device=dev); d = torch.randn(N, device=dev)
    y = torch.empty(rows, N, dtype=torch.bfloat16, device=dev)
    s = torch.rand(rows, 2, device=dev) + 0.5
    f = lambda: _lib.check(lib.hgr_linear(x.data_ptr(), rows, K, w.data_ptr(), c.data_ptr() if stats else None, d.data_ptr(), 0, None, y.data_ptr(), N, s.data_ptr() if stats else None, None, st), "lin")
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = (rows * K * 2 + rows * N * 2) / 1e9
    print(f"K={K} N={N} stats={stats}: {ms:.4f} ms  {2*rows*K*N/ms/1e9:.0f} TFLOP/s  {gb/ms*1e3:.0f} GB/s")
for K, N, stats in [(256, 768, True), (256, 768, False), (512, 768, False), (1024, 768, False), (256, 256, False), (256, 512, False), (64, 768, False)]:
    run(K, N, stats)
