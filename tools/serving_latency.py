#!/usr/bin/env python
"""Latency of ClassifierSession.run (the onnxruntime-shaped serving call of detect.py:143-145) at the batch sizes a
camera frame produces, CUDA-graph replay against the plain launch sequence.  Prints one JSON line per batch size."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "hand-gesture-recognition_b200"):
    sys.path.insert(0, str(p))
import numpy as np
import torch
import bench
from hgr_b200 import MultiTaskNet
from hgr_b200.serving import ClassifierSession


def med_ms(fn, n):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[int(len(ts) * 0.99)]


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = MultiTaskNet(21, 19, [192, 192]); bench.synthetic_weights(m); m = m.to(dev).eval()
    graph, eager = ClassifierSession(m, cuda_graph=True), ClassifierSession(m, cuda_graph=False)
    name = graph.get_inputs()[0].name
    rng = np.random.default_rng(0)
    for b in (1, 2, 8, 32):
        x = rng.standard_normal((b, 3, 192, 192), dtype=np.float32)
        a, c = graph.run(None, {name: x}), eager.run(None, {name: x})
        same = bool(np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]))
        for _ in range(20):
            graph.run(None, {name: x}); eager.run(None, {name: x})
        g50, g99 = med_ms(lambda: graph.run(None, {name: x}), 300)
        e50, e99 = med_ms(lambda: eager.run(None, {name: x}), 300)
        print(json.dumps({"batch": b, "graph_ms_p50": round(g50, 4), "graph_ms_p99": round(g99, 4),
                          "eager_ms_p50": round(e50, 4), "eager_ms_p99": round(e99, 4),
                          "outputs_identical": same}), flush=True)


main()
