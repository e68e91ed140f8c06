#!/usr/bin/env python
"""Experiment: two half-batch forwards on two streams, each persistent grid capped at half the SMs (HGR_SM_LIMIT=74),
so that HBM-bound launches of one stream overlap tensor-bound launches of the other.  Prints images/s."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "hand-gesture-recognition_b200"):
    sys.path.insert(0, str(p))
import torch
import bench
from hgr_b200 import MultiTaskNet

def main():
    B, S, K = 1024, 192, 40
    dev = torch.device("cuda", 0)
    nstreams = int(os.environ.get("NSTREAMS", "2"))
    models, xs, streams = [], [], []
    for i in range(nstreams):
        torch.manual_seed(0)
        m = MultiTaskNet(21, 19, [S, S]); bench.synthetic_weights(m); m = m.to(dev).eval(); m.return_attention = False
        models.append(m); xs.append(torch.randn(B // nstreams, 3, S, S, device=dev).to(torch.bfloat16)); streams.append(torch.cuda.Stream(dev))
    with torch.no_grad():
        for _ in range(5):
            for m, x, st in zip(models, xs, streams):
                with torch.cuda.stream(st):
                    m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams: st.wait_event(e0)
        for _ in range(K):
            for m, x, st in zip(models, xs, streams):
                with torch.cuda.stream(st):
                    m(x)
        for st in streams: torch.cuda.current_stream().wait_stream(st)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"streams {nstreams} sm_limit {os.environ.get('HGR_SM_LIMIT','-')}: {B / ms * 1e3:.0f} img/s, {ms:.3f} ms per 1024")
main()
