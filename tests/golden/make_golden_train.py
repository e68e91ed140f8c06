"""Generates tests/golden/train_step_s{64,192}.npz by running the REAL reference in training mode.

Run in the build container only:   python tests/golden/make_golden_train.py

What is pinned: reference `MultiTaskNet(21, 19, [S, S]).train()` loaded with oracle.synthetic_state_dict(seed),
one training forward on oracle.synthetic_images / oracle.synthetic_targets, the reference's own losses
(libs/loss.py: JointsMSELoss(use_target_weight=True), ClassificationLoss) combined as train.py:63-75, one
backward, one torch.optim.AdamW(lr=1e-3) step.  Stored: the three loss values, logits, a heatmap subsample, the
L2 norm of every parameter gradient, the small gradients in full, a strided subsample of the large ones, the
BatchNorm running statistics after the step and checksums of the updated parameters.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from model.multitasknet import MultiTaskNet  # noqa: E402  (the reference)
from libs.loss import ClassificationLoss, JointsMSELoss  # noqa: E402  (the reference)

from oracle import multitasknet_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    torch.set_num_threads(8)
    for size, seed, batch in [(64, 5, 4), (192, 11, 2)]:
        sd = O.synthetic_state_dict(seed)
        m = MultiTaskNet(21, 19, [size, size])
        m.load_state_dict(sd, strict=True)
        m.train()
        x = O.synthetic_images(batch, size, seed + 1)
        labels, target, weight = O.synthetic_targets(batch, size, seed=seed + 2)
        opt = torch.optim.AdamW(m.parameters(), 1e-3)
        cls, hm, _ = m(x.clone())
        cl = ClassificationLoss()(cls, labels) * 0.001
        jl = JointsMSELoss(use_target_weight=True)(hm, target, weight)
        tot = cl + jl
        opt.zero_grad()
        tot.backward()
        d = {"loss3": np.array([tot.item(), cl.item(), jl.item()]), "logits": cls.detach().numpy(),
             "heat_sub": hm.detach()[:, :, ::4, ::4].numpy()}
        for k, p in m.named_parameters():
            g = p.grad.detach()
            d["gnorm/" + k] = np.array([g.double().norm().item()])
            d["gsub/" + k] = g.flatten()[:: max(1, g.numel() // 512)].numpy() if g.numel() > 4096 else g.numpy()
        opt.step()
        for k, b in m.named_buffers():
            if "running_" in k:
                d["stat/" + k] = b.detach().numpy()
        for k, p in m.named_parameters():
            d["pnew_sum/" + k] = np.array([p.detach().double().sum().item(), p.detach().double().abs().sum().item()])
        np.savez_compressed(OUT / f"train_step_s{size}.npz", **d)
        print("wrote", size, d["loss3"])


if __name__ == "__main__":
    main()
