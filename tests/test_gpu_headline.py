"""Parity ON the configuration the headline number is quoted on (BASELINE.json configs[1]): batch 1024, 192 x 192,
bf16 in / bf16 out, return_attention=False, non-degenerate synthetic weights - the exact call bench.py times.

Round 1 gated parity at batch 8 (one tile per CTA) and top-1 on 64 samples.  At batch 1024 every persistent kernel
walks 8..125 tiles per CTA, the accumulator phase bits wrap, the zig-zag order and the CTA pairs see full waves.
The fp32 oracle (CPU) is run in chunks of 128 crops and compared with the matching slice of each HBM stage buffer.

Tolerances are the ones of tests/test_gpu_forward.py (rel-L2 <= 1.5e-2, max-abs <= 4e-2 * max|ref| per stage), i.e.
below the reference's own `.bfloat16()`-vs-fp32 error on the same weights (1.0e-2..1.6e-2).
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import multitasknet_oracle as O
from tests.helpers import report

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

REL_TOL = 1.5e-2
MAX_TOL = 4e-2


class _Acc:
    """rel-L2 / max-abs accumulated over chunks without holding the full-batch fp32 reference."""

    def __init__(self):
        self.err2 = self.ref2 = 0.0
        self.max_err = self.max_ref = 0.0
        self.nan = 0

    def add(self, got, ref):
        got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
        assert got.shape == ref.shape, (tuple(got.shape), tuple(ref.shape))
        d = got - ref
        self.err2 += float((d * d).sum())
        self.ref2 += float((ref * ref).sum())
        self.max_err = max(self.max_err, float(d.abs().max()))
        self.max_ref = max(self.max_ref, float(ref.abs().max()))
        self.nan += int(got.isnan().sum())

    def result(self, name):
        r = (self.err2 / max(self.ref2, 1e-300)) ** 0.5
        m = self.max_err / max(self.max_ref, 1e-30)
        print(f"[parity] {name}: rel_l2={r:.3e} max_abs={self.max_err:.3e} (max|ref|={self.max_ref:.3e}) nan_out={self.nan}",
              flush=True)
        return r, m


def _build(size, seed, sensitise=False):
    from hgr_b200 import MultiTaskNet
    sd = O.synthetic_state_dict(seed)
    if sensitise:
        sd = O.sensitise_class_path(sd)
    m = MultiTaskNet(21, 19, [size, size])
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def test_headline_batch1024_bf16_io_per_stage():
    size, batch, chunk = 192, 1024, 128
    m, sd = _build(size, 0)
    m.return_attention = False
    x = O.synthetic_images(batch, size, 21).to(torch.bfloat16)  # the module sees bf16, the oracle the same values
    with torch.no_grad():
        cls, hm, attn = m(x.cuda())
    torch.cuda.synchronize()
    assert attn is None and cls.dtype == hm.dtype == torch.bfloat16
    plan = m.plan_for(batch, torch.device("cuda", torch.cuda.current_device()))
    stages = ["a1", "o1", "d1", "o2", "d2", "o3"]
    if any("conv1+" in l[0] for l in plan.launch_table()):
        stages.remove("a1")  # conv1 -> conv2 -> cspelan1.cv1 runs as one kernel on a bf16 batch: a1 never reaches HBM
        # (test_stem_fused compares that kernel with the launches it replaces; HGR_STEM_FUSED=0 below runs them here)
    bufs = {n: plan.buffer(n) for n in stages}
    tok = plan.buffer("tokens").reshape(batch, -1, 256)
    acc = {n: _Acc() for n in stages + ["tokens_l3", "logits", "heatmaps"]}
    agree = conf = 0
    for c0 in range(0, batch, chunk):
        taps = {}
        cls_ref, hm_ref, _ = O.multitasknet_forward(sd, x[c0:c0 + chunk].float(), taps)
        for n in stages:
            acc[n].add(bufs[n][c0:c0 + chunk].float().permute(0, 3, 1, 2), taps[n])
        acc["tokens_l3"].add(tok[c0:c0 + chunk].float(), taps["tokens_l3"])
        acc["logits"].add(cls[c0:c0 + chunk].float(), cls_ref)
        acc["heatmaps"].add(hm[c0:c0 + chunk].float(), hm_ref)
        got = cls[c0:c0 + chunk].float().cpu()
        err = float((got - cls_ref).abs().max())
        top2 = cls_ref.topk(2, dim=1).values
        sure = (top2[:, 0] - top2[:, 1]) > 4 * err
        conf += int(sure.sum())
        agree += int(((got.argmax(1) == cls_ref.argmax(1)) & sure).sum())
        del taps
    worst_r = worst_m = 0.0
    for n in acc:
        r, mm = acc[n].result(f"headline b1024 bf16-io {n}")
        assert acc[n].nan == 0
        worst_r, worst_m = max(worst_r, r), max(worst_m, mm)
    print(f"[parity] headline b1024 top-1: {agree}/{conf} confident samples agree", flush=True)
    assert worst_r <= REL_TOL and worst_m <= MAX_TOL
    assert conf > 0 and agree == conf


def test_top1_margin_aware_2048_samples_on_the_sensitised_class_path():
    """SURVEY.md 8(c) recipe 3: class path sensitised so that the logits differ between samples (several classes,
    margins from ~0), 2048 samples, bf16 in / out.  north_star asks for >= 99.9 % top-1 agreement; a bf16 path
    cannot agree on samples whose fp32 margin is inside its own rounding error, so the figure is reported three ways:
    raw, margin-aware (all samples whose fp32 top-1/top-2 margin exceeds 4x the measured max-abs logit error must
    agree) and next to the REFERENCE's own bf16-autocast-vs-fp32 disagreement on the same samples (the noise floor:
    the same oracle graph run by the stock PyTorch operators under torch.autocast(bfloat16))."""
    size, total, chunk = 192, 2048, 256
    m, sd = _build(size, 0, sensitise=True)
    m.return_attention = False
    dev = torch.device("cuda")
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    got, ref, ref_bf16 = [], [], []
    try:
        for c0 in range(0, total, chunk):
            x = O.synthetic_images(chunk, size, 1000 + c0).to(torch.bfloat16)
            with torch.no_grad():
                cls, _, _ = m(x.cuda())
                got.append(cls.float().cpu())
                ref.append(O.multitasknet_forward(sd, x.float())[0])
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    ref_bf16.append(O.multitasknet_forward(sd_dev, x.float().to(dev))[0].float().cpu())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    got, ref, ref_bf16 = torch.cat(got), torch.cat(ref), torch.cat(ref_bf16)
    top2 = ref.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]

    def stats(out, name):
        err = float((out - ref).abs().max())
        agree = out.argmax(1) == ref.argmax(1)
        sure = margin > 4 * err
        print(f"[parity] {name}: max-abs logit error {err:.4f}, raw top-1 agreement {float(agree.float().mean()):.4f} "
              f"({int(agree.sum())}/{len(agree)}), confident {int((agree & sure).sum())}/{int(sure.sum())}", flush=True)
        return err, float(agree.float().mean()), int(sure.sum()), int((agree & sure).sum())

    print(f"[parity] sensitised recipe: {len(set(ref.argmax(1).tolist()))} distinct classes, logit batch-std "
          f"{float(ref.std(0).mean()):.3f}, margin quantiles 1%/10%/50% = "
          f"{float(margin.quantile(0.01)):.4f}/{float(margin.quantile(0.1)):.4f}/{float(margin.quantile(0.5)):.4f}",
          flush=True)
    e1, raw1, n1, ok1 = stats(got, "this path (bf16 io) vs fp32 oracle")
    e0, raw0, n0, ok0 = stats(ref_bf16, "reference graph under bf16 autocast (stock operators) vs fp32 oracle")
    assert len(set(ref.argmax(1).tolist())) >= 3, "degenerate recipe: the class path does not vary"
    assert n1 >= total // 2 and ok1 == n1                 # every confident sample agrees
    assert 1.0 - raw1 <= max(2.0 * (1.0 - raw0), 0.004)   # raw disagreement no worse than ~the reference's own floor
    assert e1 <= max(2.0 * e0, 0.05)


_WORKER = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r})
from oracle import multitasknet_oracle as O
from hgr_b200 import MultiTaskNet
m = MultiTaskNet(21, 19, [192, 192]); m.load_state_dict(O.synthetic_state_dict(0), strict=True); m = m.cuda().eval()
m.return_attention = False
x = O.synthetic_images({batch}, 192, 31).to(torch.bfloat16).cuda()
with torch.no_grad():
    cls, hm, _ = m(x)
torch.cuda.synchronize()
plan = m.plan_for({batch}, torch.device("cuda", torch.cuda.current_device()))
names = [l[0] for l in plan.launch_table()]
np.savez({out!r}, cls=cls.float().cpu().numpy(), hm=hm[:, :, ::3, ::3].float().cpu().numpy(),
         o3=plan.buffer("o3").float().cpu().numpy(), launches=np.array(len(names)))
"""


@pytest.mark.parametrize("switch", ["HGR_STEM_FUSED=0", "HGR_GELAN_TAIL=0", "HGR_CONV_CHAIN=0", "HGR_CHAIN_HALO=0",
                                    "HGR_VIT_FUSED=0", "HGR_CLUSTER=0", "HGR_ZIGZAG=0", "HGR_ATTN_TC=0", "HGR_POSE_TC=0"])
def test_toggled_launch_paths_agree_with_the_default_path(switch, tmp_path):
    """Every run-time switch selects a different kernel or tile order for the same arithmetic; the library reads them
    once per process, so each variant runs in its own interpreter.  Batch 256 gives every persistent kernel several
    tiles per CTA (o3: 288 m-tiles x 2 n-tiles over 148 CTAs).  The variants must agree with the default path to within
    the bf16 rounding of the stages they re-associate (chained vs. separate launches round at the same places, the
    attention kernels differ in summation order), far below the tolerance against the oracle."""
    batch = 256
    outs = {}
    for tag, extra in (("default", {}), ("variant", dict([switch.split("=")]))):
        out = tmp_path / f"{tag}.npz"
        env = dict(os.environ, **extra)
        code = _WORKER.format(root=str(ROOT), pkg=str(ROOT / "hand-gesture-recognition_b200"), batch=batch, out=str(out))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = np.load(out)
    a, b = outs["default"], outs["variant"]
    print(f"[parity] {switch}: launches {int(a['launches'])} -> {int(b['launches'])}", flush=True)
    # two bf16 paths that differ in ONE rounding anywhere upstream decorrelate layer by layer and end up about sqrt(2) x
    # the rounding share of their own error against fp32 apart (o3: 9.8e-3 against the oracle), so the bar between two
    # variants is 1e-2 - still below the 1.5e-2 bar against the oracle
    for key, tol in (("o3", 1e-2), ("cls", 1e-2), ("hm", 1e-2)):
        r, _ = report(f"{switch} {key} vs default path", torch.from_numpy(b[key]), torch.from_numpy(a[key]))
        assert r <= tol
