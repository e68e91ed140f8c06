#!/usr/bin/env python
"""Experiment: the batch-1024 forward replayed from a CUDA graph (torch.cuda.CUDAGraph around MultiTaskNet.forward)
against the plain stream launch.  Prints images/s for both."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "hand-gesture-recognition_b200"):
    sys.path.insert(0, str(p))
import torch
import bench
from hgr_b200 import MultiTaskNet


def timed(fn, k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def main():
    B, S, K = 1024, 192, 40
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = MultiTaskNet(21, 19, [S, S]); bench.synthetic_weights(m); m = m.to(dev).eval(); m.return_attention = False
    x = torch.randn(B, 3, S, S, device=dev).to(torch.bfloat16)
    with torch.no_grad():
        for _ in range(5):
            m(x)
        ms_plain = timed(lambda: m(x), K)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            m(x)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = m(x)
        for _ in range(3):
            g.replay()
        ms_graph = timed(g.replay, K)
        ms_plain2 = timed(lambda: m(x), K)
    ref = m(x)
    g.replay(); torch.cuda.synchronize()
    same = torch.equal(ref[0], out[0]) and torch.equal(ref[1], out[1])
    print(f"plain {B / ms_plain * 1e3:.0f} img/s ({ms_plain:.3f} ms)  graph {B / ms_graph * 1e3:.0f} img/s ({ms_graph:.3f} ms)  "
          f"plain again {B / ms_plain2 * 1e3:.0f} img/s ({ms_plain2:.3f} ms)  outputs identical: {same}")


main()
