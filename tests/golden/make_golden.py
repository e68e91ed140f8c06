"""Generates tests/golden/*.npz by running the REAL reference implementation.

Run in the build container only (it imports /root/reference, which does not
exist on the GPU box):   python tests/golden/make_golden.py

What is pinned
  multitasknet_s{192,256}_seed{S}.npz
      reference `model.multitasknet.MultiTaskNet(21, 19, [S, S]).eval()` loaded
      (strict=True) with oracle.synthetic_state_dict(seed) and fed
      oracle.synthetic_images(batch, S, seed+1): logits (full), heatmaps and
      last-layer attention (strided subsample + fp64 checksums), per-stage
      mean / std / abs-sum of the backbone and transformer intermediates
      captured with forward hooks on the reference module.
  default_init_s192.npz
      the literal BASELINE.json config #1 recipe: torch.manual_seed(0), default
      init, x = randn(4, 3, 192, 192).
  get_max_preds.npz
      libs.utils.get_max_preds on the crafted maps of cases.heatmap_cases():
      ties, all-negative, zeros, NaN, +-inf, 48x48, 64x64, non-square.
  crop_normalize.npz
      the detect.py:106-112 arithmetic on every uint8 value per channel.
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
from libs.utils import get_max_preds  # noqa: E402  (the reference)

from oracle import multitasknet_oracle as O  # noqa: E402
from tests.golden.cases import crop_image, heatmap_cases  # noqa: E402

OUT = Path(__file__).resolve().parent


def stats(t):
    t = t.double()
    return np.array([t.mean().item(), t.std().item(), t.abs().sum().item()])


def run_reference(model, x):
    """Forward with hooks on the modules whose outputs the per-stage tests compare."""
    taps = {}
    hooks = []

    def hook(name):
        def fn(_m, _i, out):
            taps[name] = (out[0] if isinstance(out, tuple) else out).detach().clone()  # ViT adds pe in place later
        return fn

    enc = model.encoder
    for name, mod in [("a1", enc.conv1), ("a2", enc.conv2), ("o1", enc.cspelan1), ("d1", enc.down1),
                      ("o2", enc.cspelan2), ("d2", enc.down2), ("o3", enc.cspelan3), ("proj", model.proj)]:
        hooks.append(mod.register_forward_hook(hook(name)))
    with torch.no_grad():
        cls, hm, attn = model(x.clone())
    for h in hooks:
        h.remove()
    return cls, hm, attn, taps


def pack_outputs(cls, hm, attn, taps):
    d = {"logits": cls.numpy(), "heat_sub": hm[:, :, ::4, ::4].numpy(), "heat_stats": stats(hm),
         "attn_sub": attn[:, :, ::8, ::8].numpy(), "attn_stats": stats(attn),
         "heat_row": hm[0, 0].numpy()}
    for k, v in taps.items():
        d["stats_" + k] = stats(v)
    return d


def main():
    torch.set_num_threads(8)
    for size, seed, batch in [(192, 0, 4), (192, 7, 2), (256, 3, 2)]:
        sd = O.synthetic_state_dict(seed)
        m = MultiTaskNet(21, 19, [size, size]).eval()
        m.load_state_dict(sd, strict=True)
        x = O.synthetic_images(batch, size, seed + 1)
        d = pack_outputs(*run_reference(m, x))
        d["x_stats"] = stats(x)
        np.savez_compressed(OUT / f"multitasknet_s{size}_seed{seed}.npz", **d)
        print("wrote", size, seed, d["logits"].shape, d["heat_sub"].shape)

    torch.manual_seed(0)
    m = MultiTaskNet(21, 19, [192, 192]).eval()
    x = torch.randn(4, 3, 192, 192)
    d = pack_outputs(*run_reference(m, x))
    # default init: keep the conv1 weight checksum so the test can prove it built the same weights
    d["conv1_w_sum"] = np.array([m.state_dict()["encoder.conv1.conv.weight"].double().sum().item()])
    np.savez_compressed(OUT / "default_init_s192.npz", **d)

    cases = heatmap_cases()
    out = {}
    for name, maps in cases.items():
        p, v = get_max_preds(maps)
        out["preds_" + name], out["maxvals_" + name] = p, v
    np.savez_compressed(OUT / "get_max_preds.npz", **out)

    # detect.py:106-112 on an image holding every byte value in every channel
    img = crop_image()
    im = img.transpose((2, 0, 1)).astype(np.float32)
    im /= 255
    mean = np.array([0.485, 0.456, 0.406], dtype=np.float32)
    std = np.array([0.229, 0.224, 0.225], dtype=np.float32)
    im = (im - mean.reshape(3, 1, 1)) / std.reshape(3, 1, 1)
    np.savez_compressed(OUT / "crop_normalize.npz", out=np.ascontiguousarray(np.expand_dims(im, 0)))
    print("done")


if __name__ == "__main__":
    main()
