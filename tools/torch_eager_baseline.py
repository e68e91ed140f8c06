#!/usr/bin/env python
"""Secondary baseline: the reference's algorithm run by stock PyTorch on the SAME B200 -
cuDNN convolutions, cuBLAS GEMMs, ATen element-wise kernels - i.e. the path a user of the reference gets today
on this GPU.  /root/reference does not exist on the GPU box, so the oracle's restatement (same ATen operators,
pinned against the real reference) stands in for it.  Measurement only: nothing here is product code.

    python tools/torch_eager_baseline.py [--batch 1024] [--size 192] [--iters 10]
prints one JSON line per variant (fp32, TF32, bf16 autocast, bf16 autocast + channels_last).
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import multitasknet_oracle as O  # noqa: E402


def run(sd, x, iters, autocast, channels_last):
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
        sd = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            feat = O.gelan_net(sd, x)
            feat = torch.nn.functional.conv2d(feat, sd["proj.weight"])
            return O.vit(sd, feat)

    for _ in range(3):
        out = fwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fwd()
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(out[0].float()).all()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--size", type=int, default=192)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda")
    sd = {k: v.to(dev) for k, v in O.synthetic_state_dict(0).items()}
    x = torch.randn(a.batch, 3, a.size, a.size, device=dev)
    torch.backends.cudnn.benchmark = True
    for name, tf32, ac, cl in [("fp32", False, False, False), ("tf32", True, False, False),
                               ("bf16_autocast", True, True, False), ("bf16_autocast_channels_last", True, True, True)]:
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        ms = run(sd, x, a.iters, ac, cl)
        print(json.dumps({"baseline": "stock PyTorch eager on the same B200 (oracle restatement: cuDNN / cuBLAS / ATen)",
                          "variant": name, "batch": a.batch, "image_size": a.size, "ms_per_step": ms,
                          "images_per_s": a.batch / (ms * 1e-3), "torch": torch.__version__,
                          "cudnn": torch.backends.cudnn.version()}), flush=True)


if __name__ == "__main__":
    main()
