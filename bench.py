#!/usr/bin/env python
"""Benchmark of the B200-native MultiTaskNet forward path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--batch B] [--size S]

One "step" is one forward pass of MultiTaskNet (GELAN backbone -> ViT -> class
logits + pose heatmaps) over one batch of synthetic hand crops.  Metric (from
BASELINE.json): hand-crop images/s, bf16, 192x192, batch 1024 per GPU.

  value     whole-job images/s with the bf16 NCHW input already resident in HBM
            (CUDA events around K steps, max over ranks)
  e2e       the same metric through the host-facing HandPipeline: uint8 crops in
            pinned HOST memory -> H2D -> crop normalise -> forward -> keypoint
            decode -> logits/keypoints D2H, every step, copies inside the timed
            region (two lanes overlap step i+1's copy with step i's kernels)
  roofline  the tcgen05 kernels of a step: algorithmic FLOPs / CUDA-event time of
            those launches, measured live in a profiling pass, against
            MEASURED_PEAKS.json; `per_kernel` lists every launch with its own bound
  roofline_memory  the HBM-bound tail (get_max_preds, crop normalise, fused crop
            warp, conv1): achieved GB/s of their algorithmic bytes against the
            measured copy bandwidth
  parity    a 32-crop slice of the timed batch against the fp32 oracle, checked
            OUTSIDE the timed region (rel-L2 of logits and heatmaps)
  cpu_baseline  the reference forward on the box's host cores on a bounded sample
            (rank 0, N = 1 only): the real reference modules when oracle/_ref is
            staged (`kind: "reference"`), else the oracle's restatement (`"port"`)
  gpu_eager_baseline  the same reference graph run by stock PyTorch (cuDNN / cuBLAS,
            bf16 autocast, channels_last) on the same GPU in the same run
  train     the data-parallel training step of BASELINE.json configs[4] (batch 32
            per GPU): step time and the exposed all-reduce, at the same N

`--impl reference` times the CPU path alone.  `--total-batch 8192` is
BASELINE.json configs[2]: ONE batch sharded contiguously over the ranks
(strong scaling).  Under torchrun every rank drives one GPU with its own shard
(no data-path collective); rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "hand-gesture-recognition_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GFLOP_PER_IMG = {192: 4.374, 256: 7.891}  # BASELINE.md section 2 (2 x MAC, unpadded)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="crops per GPU per step")
    ap.add_argument("--size", type=int, default=192)
    ap.add_argument("--total-batch", type=int, default=0,
                    help="BASELINE.json configs[2]: one batch of this many crops sharded over the ranks (strong scaling)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="N > 1: do not bind the rank to the CPU socket of its GPU (A/B of the e2e path)")
    ap.add_argument("--train-graph", action="store_true",
                    help="--workload train: replay forward + loss + backward from a CUDA graph")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the memory-tail roofline, the GPU eager baseline, the parity slice and the train record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="", help="write the per-launch table (json) here")
    ap.add_argument("--workload", default="forward", choices=["forward", "train"],
                    help="forward = the headline metric (BASELINE.json configs[1]); train = the data-parallel "
                         "training step of configs[4] (batch 32 per GPU unless --batch is given)")
    return ap.parse_args()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


def ncu_traffic_per_launch():
    """DRAM bytes per launch of the tcgen05 GEMM kernel from the committed `ncu --set full` capture of one step
    (profiles/*_step_ncu_full.csv: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    import csv
    files = sorted((ROOT / "profiles").glob("*_step_ncu_full.csv"))
    if not files:
        return None, None
    rows = [r for r in csv.DictReader(files[-1].open()) if any(k in r["Kernel Name"] for k in ("gemm_kernel", "halo", "vit_block", "stem_chain", "conv_chain", "stem_umma", "gelan_tail"))]
    if not rows:
        return None, None
    rd = [k for k in rows[0] if k.startswith("dram__bytes_read.sum")][0]
    wr = [k for k in rows[0] if k.startswith("dram__bytes_write.sum")][0]
    units = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    u_rd, u_wr = (units[k.split("[")[1].rstrip("]")] for k in (rd, wr))  # ncu picks the unit per column
    total = sum(float(r[rd]) * u_rd + float(r[wr]) * u_wr for r in rows)
    return total / len(rows), files[-1].name


def synthetic_weights(model, seed=0):
    """He-scaled conv/linear weights and non-trivial BN statistics, so activations stay O(1) through the net."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        if k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif k.endswith("running_mean") or k.endswith("bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.2)
        elif v.dim() == 1:
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif v.dim() >= 2 and k != "decoder.cls_token":
            fan_in = v[0].numel()
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5)
    model.load_state_dict(sd)


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """Samples taken from here on belong to the timed region."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = self.rows[self.first:] or self.rows[-3:]
        clocks = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        smax = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        busy = [c for c in clocks if smax and c > 0.3 * smax[0]] or clocks
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": smax[0] if smax else None,
                "reasons": reasons, "samples": len(clocks)}


def reference_forward_fn(size, device="cpu"):
    """The reference's forward as a callable x -> outputs, with the synthetic weights the oracle defines: the REAL
    reference modules when oracle/_ref is staged (oracle/build_ref.py), else the oracle's restatement of them.
    Checker / baseline only - nothing here is on the product path.  Returns (fn, kind, description)."""
    import torch
    from oracle import build_ref
    from oracle import multitasknet_oracle as O
    sd = O.synthetic_state_dict(0)
    staged = build_ref.load()
    if staged is not None:
        net = staged[0](21, 19, [size, size])
        net.load_state_dict(sd, strict=True)
        net = net.eval().to(device)

        def fn(x):
            with torch.no_grad():
                return net(x)
        return fn, "reference", "unmodified reference modules (oracle/_ref: model/multitasknet.py + gelan.py + transformer.py)"
    sd = {k: v.to(device) for k, v in sd.items()}

    def fn(x):
        with torch.no_grad():
            return O.multitasknet_forward(sd, x)
    return fn, "port", "oracle restatement of the reference forward (same ATen operators)"


def cpu_forward_rate(size, batch=32, budget_s=12.0, min_iters=3):
    """The reference's fp32 forward on all host cores: (img/s, cores, kind, description)."""
    import torch
    from oracle import multitasknet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, kind, what = reference_forward_fn(size)
    x = O.synthetic_images(batch, size, 1)
    fn(x)  # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_iters or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        fn(x)
        times.append(time.perf_counter() - t0)
        if len(times) >= 64:
            break
    med = statistics.median(times)
    return batch / med, cores, kind, f"{what}, fp32, batch {batch} x {len(times)} iterations at {size}x{size}, median"


def gpu_eager_rate(size, batch, dev, iters=6):
    """The same reference graph on THIS GPU through stock PyTorch (cuDNN / cuBLAS dispatch), bf16 autocast and
    channels_last: the 'existing Blackwell path' of BASELINE.md section 4, a reported baseline."""
    import torch
    from oracle import multitasknet_oracle as O
    fn, kind, what = reference_forward_fn(size, dev)
    x = O.synthetic_images(32, size, 1).to(dev)
    x = x.repeat((batch + 31) // 32, 1, 1, 1)[:batch].contiguous(memory_format=torch.channels_last)
    best = None
    for bsz in (batch, batch // 4):  # the eager graph keeps more activations alive than the fused path
        try:
            xb = x[:bsz]
            with torch.autocast("cuda", dtype=torch.bfloat16):
                for _ in range(2):
                    fn(xb)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    fn(xb)
                e1.record()
            torch.cuda.synchronize(dev)
            best = (bsz * iters / (e0.elapsed_time(e1) * 1e-3), bsz)
            break
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
    if best is None:
        return None
    return {"value": best[0], "unit": "images/s", "batch": best[1], "kind": kind,
            "what": f"{what} on the same GPU: stock PyTorch eager, torch.autocast(bfloat16), channels_last input, "
                    f"{iters} iterations, CUDA events"}


def forward_other_size(size, batch, dev, pk, steps=8):
    """BASELINE.json configs[3] in the default line: the same forward at another image side (256 x 256), device
    resident, bf16 in / out, CUDA events over `steps` forwards after 3 warm-up forwards."""
    import torch
    from hgr_b200 import MultiTaskNet
    torch.manual_seed(0)
    model = MultiTaskNet(21, 19, [size, size])
    synthetic_weights(model)
    model = model.to(dev).eval()
    model.return_attention = False
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(batch, 3, size, size, generator=g, device=dev).bfloat16()
    with torch.no_grad():
        for _ in range(3):
            out = model(x)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = model(x)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    ok = bool(torch.isfinite(out[0].float()).all() and torch.isfinite(out[1].float()).all())
    gf = GFLOP_PER_IMG.get(size)
    value = batch / (ms * 1e-3)
    return {"metric": "hand-crop images/s, MultiTaskNet forward (BASELINE.json configs[3])", "image_size": size,
            "batch": batch, "value": value, "unit": "images/s", "ms_per_step": ms, "steps": steps, "finite": ok,
            "net_tensor_frac_of_sustained": value * gf * 1e9 / (pk["bf16_sustained"] * 1e12) if gf else None,
            "net_tensor_frac_of_burst": value * gf * 1e9 / (pk["bf16_burst"] * 1e12) if gf else None}


def memory_tail_roofline(model, B, S, dev, hbm_gbs, conv1_row):
    """Achieved GB/s of the HBM-bound kernels around the forward (north_star: 'achieved HBM GB/s for the memory-bound
    kernels'): algorithmic bytes / CUDA-event time over 20 launches that alternate between two buffer sets larger
    than L2 together."""
    import numpy as np
    import torch
    from hgr_b200 import _lib
    from hgr_b200.ops import box_to_affine, invert_affine
    lib = _lib.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    J, hs = model.num_joints, S // 4
    out = []

    def timed(name, launch, nbytes, what, reps=20):
        for i in range(3):
            launch(i & 1)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            launch(i & 1)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "ms": ms, "bytes": nbytes, "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s",
                    "frac": gbs / hbm_gbs, "what": what})

    g = torch.Generator(device=dev).manual_seed(3)
    heat = [torch.randn(B, J, hs, hs, generator=g, device=dev) for _ in range(2)]
    preds = torch.empty(B, J, 2, device=dev)
    maxv = torch.empty(B, J, 1, device=dev)
    timed("max_preds_kernel (libs/utils.py:4-32)",
          lambda k: _lib.check(lib.hgr_get_max_preds(heat[k].data_ptr(), _lib.F32, B, J, hs, hs, preds.data_ptr(),
                                                     maxv.data_ptr(), st), "hgr_get_max_preds"),
          B * J * (hs * hs * 4 + 12), "fp32 heatmaps in, (x, y, maxval) out")
    del heat
    crops = [torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    xn = [torch.empty(B, 3, S, S, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    timed("crop_normalize_kernel (detect.py:106-112)",
          lambda k: _lib.check(lib.hgr_crop_normalize(crops[k].data_ptr(), xn[k].data_ptr(), _lib.BF16, B, S, S, st),
                               "hgr_crop_normalize"),
          B * S * S * 3 * (1 + 2), "uint8 HWC crops in, bf16 CHW out")
    # fused warp: B boxes of ~S x S pixels spread over 64 camera frames (640 x 480), one crop each
    nf, hf, wf = 64, 480, 640
    frames = [torch.randint(0, 256, (nf, hf, wf, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    rng = np.random.default_rng(0)
    x0 = rng.integers(0, wf - S, B)
    y0 = rng.integers(0, hf - S, B)
    inv = np.stack([invert_affine(box_to_affine((a, b, a + S, b + S), S)) for a, b in zip(x0, y0)])
    d_inv = torch.from_numpy(np.ascontiguousarray(inv)).to(dev)
    d_idx = torch.from_numpy((np.arange(B) % nf).astype(np.int32)).to(dev)
    timed("crop_warp_normalize_kernel (detect.py:92-117)",
          lambda k: _lib.check(lib.hgr_crop_warp_normalize(frames[k].data_ptr(), nf, hf, wf, d_idx.data_ptr(),
                                                           d_inv.data_ptr(), B, S, xn[k].data_ptr(), _lib.BF16, st),
                               "hgr_crop_warp_normalize"),
          B * S * S * 3 * (1 + 2), "S x S source pixels of a uint8 frame per crop in, bf16 CHW out")
    if conv1_row is not None:
        nb = B * (3 * S * S * 2 + 64 * (S // 2) * (S // 2) * 2)
        out.append({"kernel": "conv1_kernel (model/gelan.py:155, K = 27)", "ms": conv1_row["ms"], "bytes": nb,
                    "achieved": nb / (conv1_row["ms"] * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                    "frac": nb / (conv1_row["ms"] * 1e-3) / 1e9 / hbm_gbs,
                    "what": "bf16 NCHW input in, bf16 NHWC 64-channel map out; timed inside the forward"})
    else:
        # the default forward runs conv1 inside the fused stem kernel (a1 never reaches HBM); the stand-alone kernel
        # still serves fp32 input batches, other image sides and the training step
        del crops, frames
        wk = torch.zeros(64, 32, device=dev)
        wk[:, :27] = torch.randn(64, 27, generator=g, device=dev) * 0.27
        wk = wk.bfloat16()
        sh = torch.randn(64, generator=g, device=dev) * 0.3
        a1 = [torch.empty(B, S // 2, S // 2, 64, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        timed("conv1_kernel (model/gelan.py:155, K = 27; stand-alone launch, fused into the stem kernel in the forward)",
              lambda k: _lib.check(lib.hgr_conv1(xn[k].data_ptr(), _lib.BF16, B, S, wk.data_ptr(), sh.data_ptr(),
                                                 a1[k].data_ptr(), st), "hgr_conv1"),
              B * (3 * S * S * 2 + 64 * (S // 2) * (S // 2) * 2), "bf16 NCHW input in, bf16 NHWC 64-channel map out")
    return out


def parity_slice(model, x, out, n=32):
    """The first n crops of the timed batch against the fp32 oracle on the model's own weights (outside the timed
    region).  The oracle sees the same bf16-rounded input values the kernels saw."""
    import torch
    from oracle import multitasknet_oracle as O
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    cls_ref, hm_ref, _ = O.multitasknet_forward(sd, x[:n].float().cpu())

    def rel(a, b):
        return float((a.float().cpu() - b).norm() / b.norm())

    err = float((out[0][:n].float().cpu() - cls_ref).abs().max())
    top2 = cls_ref.topk(2, dim=1).values
    sure = (top2[:, 0] - top2[:, 1]) > 4 * err
    agree = out[0][:n].float().cpu().argmax(1) == cls_ref.argmax(1)
    return {"crops": n, "checker": "fp32 oracle on the host, same weights and inputs, outside the timed region",
            "logits_rel_l2": rel(out[0][:n], cls_ref), "heatmaps_rel_l2": rel(out[1][:n], hm_ref),
            "logits_max_abs": err, "top1_agree_confident": f"{int((agree & sure).sum())}/{int(sure.sum())}",
            "tolerance_rel_l2": 1.5e-2}


def train_record(S, dev, rank, world, local_rank, barrier, steps=10):
    """BASELINE.json configs[4] at the same N: step time of DataParallelTrainer (batch 32 per GPU) and the time of
    its one gradient all-reduce, which runs after the backward and is therefore fully exposed."""
    import torch
    import torch.distributed as dist
    from hgr_b200 import DataParallelTrainer, MultiTaskNet
    from hgr_b200.sharding import max_over_ranks
    from hgr_b200.training import allreduce_sum_
    B = 32
    torch.manual_seed(0)
    model = MultiTaskNet(21, 19, [S, S])
    synthetic_weights(model)
    model = model.to(dev).train()
    tr = DataParallelTrainer(model, lr=1e-4, cuda_graph=True)  # forward + loss + backward replayed from a CUDA graph
    g = torch.Generator(device=dev).manual_seed(11 + rank)
    x = torch.randn(B, 3, S, S, generator=g, device=dev)
    labels = torch.randint(0, 19, (B,), generator=g, device=dev)
    target = torch.rand(B, 21, S // 4, S // 4, generator=g, device=dev)
    weight = torch.ones(B, 21, 1, device=dev)
    for _ in range(3):
        loss = tr.step(x, labels, target, weight)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.step(x, labels, target, weight)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev) / steps
    ar = 0.0
    if world > 1:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(steps):
            allreduce_sum_(tr.state.grads)
        a1.record()
        barrier()
        ar = max_over_ranks(a0.elapsed_time(a1), dev) / steps
    ok = bool(torch.isfinite(loss).all())
    return {"metric": "training images/s, MultiTaskNet fwd+bwd+allreduce+AdamW (BASELINE.json configs[4])",
            "value": world * B / (ms * 1e-3), "unit": "images/s", "batch_per_gpu": B, "ms_per_step": ms,
            "allreduce_ms_exposed": ar, "allreduce_bytes": tr.state.numel * 4, "steps": steps, "finite_loss": ok,
            "cuda_graph": "forward + loss + backward replayed from one captured graph; all-reduce and AdamW are plain launches"}


def workload(B, S):
    return (f"MultiTaskNet forward bf16, batch {B} per GPU, 3x{S}x{S}, 19 classes, 21 keypoints "
            "(BASELINE.json configs[1]); attention map not materialised")


def run_reference_arm(args, rank, world):
    """CPU arm: the reference's algorithm (oracle port, same ATen operators) on the host cores; rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import multitasknet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = 32
    fn, kind, what = reference_forward_fn(args.size)
    x = O.synthetic_images(batch, args.size, 1)
    for _ in range(max(1, min(args.warmup, 3))):
        fn(x)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(x)
    dt = time.perf_counter() - t0
    v = batch * steps / dt
    sample = f"{what}, fp32 on the host cores, batch {batch} x {steps} steps at {args.size}x{args.size}"
    print(json.dumps({
        "impl": "reference", "metric": "hand-crop images/s, MultiTaskNet forward", "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.batch, args.size), "batch_per_gpu": args.batch, "image_size": args.size,
                   "sample": f"each step is a bounded sample of the workload: {batch} of the {args.batch} crops, "
                             "fp32 on the host cores (the reference's own CPU path, BASELINE.json configs[0])"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def run_train(args, rank, world, local_rank):
    """BASELINE.json configs[4]: fwd + bwd (CE + heatmap MSE) + gradient all-reduce + AdamW, batch 32 per GPU."""
    import torch
    import torch.distributed as dist
    from hgr_b200 import DataParallelTrainer, MultiTaskNet
    from hgr_b200.sharding import max_over_ranks

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from hgr_b200.sharding import bind_host_to_device_node
    numa_node = bind_host_to_device_node(local_rank) if world > 1 and not args.no_numa_bind else None  # pinned buffers on the GPU's socket
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    B = args.batch if args.batch != 1024 else 32
    S, K, W = args.size, args.steps, max(args.warmup, 3)
    torch.manual_seed(0)
    model = MultiTaskNet(21, 19, [S, S])
    synthetic_weights(model)
    model = model.to(dev).train()
    overlap = os.environ.get("HGR_TRAIN_OVERLAP", "0") == "1"  # A/B: 1 = bucketed all-reduce under the backward (measured slower)
    tr = DataParallelTrainer(model, lr=1e-4, cuda_graph=args.train_graph, overlap_allreduce=overlap)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x = torch.randn(B, 3, S, S, generator=g, device=dev)
    # synthetic supervision of the shape train.py feeds: class labels, Gaussian-blob target heatmaps (sigma 2,
    # libs/load.py:148-206), visibility weights
    labels = torch.randint(0, 19, (B,), generator=g, device=dev)
    hs = S // 4
    cy = torch.rand(B, 21, 1, 1, generator=g, device=dev) * hs
    cx = torch.rand(B, 21, 1, 1, generator=g, device=dev) * hs
    yy = torch.arange(hs, device=dev).view(1, 1, hs, 1).float()
    xx = torch.arange(hs, device=dev).view(1, 1, 1, hs).float()
    target = torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 2.0 ** 2)).contiguous()
    weight = (torch.rand(B, 21, 1, generator=g, device=dev) > 0.2).float()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        loss = tr.step(x, labels, target, weight)
    barrier()
    if rank == 0:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        loss = tr.step(x, labels, target, weight)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    lv = [float(v) for v in loss.cpu()]
    assert all(v == v for v in lv), "non-finite loss"
    if rank == 0:
        gf = GFLOP_PER_IMG.get(S)
        pk = peaks()
        value = world * B * K / (ms * 1e-3)
        print(json.dumps({
            "metric": "training images/s, MultiTaskNet fwd+bwd+allreduce+AdamW", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"MultiTaskNet training step bf16 (fp32 master weights), batch {B} per GPU, 3x{S}x{S}, "
                                   "loss 0.001*CE + JointsMSE, AdamW, NCCL all-reduce of the flat 7.4M-element gradient "
                                   "block (BASELINE.json configs[4])" + ("; forward + loss + backward replayed from a CUDA graph" if args.train_graph else ""),
                       "batch_per_gpu": B, "image_size": S, "allreduce_overlapped_with_backward": tr.overlap,
                       "train_gflop_per_image_convention_3x_forward": 3 * gf if gf else None,
                       "tensor_frac_of_sustained": value / world * 3 * gf * 1e9 / (pk["bf16_sustained"] * 1e12) if gf else None},
            "clocks": clocks, "final_loss": lv}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.workload == "train":
        run_train(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    from hgr_b200 import HandPipeline, MultiTaskNet
    from hgr_b200.sharding import max_over_ranks, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MultiTaskNet path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from hgr_b200.sharding import bind_host_to_device_node
    numa_node = bind_host_to_device_node(local_rank) if world > 1 and not args.no_numa_bind else None  # pinned buffers on the GPU's socket
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    B, S, K, W = args.batch, args.size, args.steps, max(args.warmup, 3)
    strong = args.total_batch > 0
    if strong:
        # BASELINE.json configs[2]: ONE batch, this rank forwards its contiguous shard of it
        lo, hi = shard_range(args.total_batch, rank, world)
        B = hi - lo
    total_per_step = args.total_batch if strong else world * B
    torch.manual_seed(0)
    model = MultiTaskNet(21, 19, [S, S])
    synthetic_weights(model)
    model = model.to(dev).eval()
    model.return_attention = False
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x = torch.randn(B, 3, S, S, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)

    with torch.no_grad():
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()  # nvidia-smi needs ~0.2 s to deliver its first sample: start before the warm-up
        for _ in range(W):
            out = model(x)
        barrier()
        if rank == 0:
            sampler.mark()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(K):
            out = model(x)
        ev1.record()
        barrier()
        ms_total = max_over_ranks(ev0.elapsed_time(ev1), dev)
    clocks = sampler.stop() if rank == 0 else None
    value = total_per_step * K / (ms_total * 1e-3)
    assert torch.isfinite(out[0].float()).all() and torch.isfinite(out[1].float()).all(), "non-finite outputs"
    extras = rank == 0 and not args.no_extras
    parity = parity_slice(model, x, out) if extras else None
    plan = model.plan_for(B, dev)
    launches_per_step = plan.launches()

    # ---- end to end through the host-facing pipeline --------------------------------------
    e2e = None
    if not args.no_e2e:
        pipe = HandPipeline(model, B, torch.bfloat16, lanes=2)
        crops = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8).pin_memory()
        for _ in range(3):
            pipe.collect(pipe.submit(crops))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pending = []
        for _ in range(K):
            pending.append(pipe.submit(crops, after=e0 if len(pending) < 2 else None))
            if len(pending) == 2:
                pipe.collect(pending.pop(0))
        while pending:
            res = pipe.collect(pending.pop(0))
        for ln in pipe.lanes:
            torch.cuda.current_stream().wait_stream(ln.stream)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1), dev)
        assert torch.isfinite(res[0]).all()
        e2e = {"value": total_per_step * K / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": ms_e2e / K,
               "path": "HandPipeline: pinned uint8 crops -> H2D -> crop_normalize -> forward -> get_max_preds -> "
                       "logits+keypoints D2H, 2 lanes",
               "host_numa_node": numa_node}
        launches_e2e = pipe.launches_per_batch
        del pipe

    # ---- per-launch profile: roofline of the dominant kernel -------------------------------
    roofline, table = None, []
    pk = peaks()
    if rank == 0:
        logits = torch.empty(B, 19, dtype=torch.bfloat16, device=dev)
        heat = torch.empty(B, 21, S // 4, S // 4, dtype=torch.bfloat16, device=dev)
        info = plan.launch_table()
        # 12 event-timed forwards back to back, the last 6 kept: the per-launch times then come from the same
        # power-capped clock state as the timed loop above (a single cold pass runs ~8 % faster than the loop)
        runs = [plan.profile(x, logits, heat) for _ in range(12)][6:]
        ms = [statistics.median(r[i] for r in runs) for i in range(len(info))]
        step_ms = sum(ms)
        for (name, kind, fl, by), t in zip(info, ms):
            table.append({"launch": name, "kind": ["tcgen05_gemm", "mma_sync", "memory", "tcgen05_fused"][kind], "ms": t,
                          "share": t / step_ms, "tflops": fl / (t * 1e-3) / 1e12 if t > 0 else None,
                          "gbs": by / (t * 1e-3) / 1e9 if t > 0 else None})
        gem = [(fl, t) for (name, kind, fl, by), t in zip(info, ms) if kind == 0]
        g_fl, g_ms = sum(f for f, _ in gem), sum(t for _, t in gem)
        ach = g_fl / (g_ms * 1e-3) / 1e12
        traffic, traffic_src = ncu_traffic_per_launch() if (B, S) == (1024, 192) else (None, None)
        g_by = sum(by for (name, kind, fl, by) in info if kind == 0)
        roofline = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_sustained"], "traffic": traffic,
                    "traffic_source": (f"profiles/{traffic_src}: dram__bytes_read.sum + dram__bytes_write.sum, bytes per "
                                       "launch averaged over the GEMM launches of one step") if traffic else None,
                    "algorithmic_bytes_per_launch_avg": g_by / len(gem),
                    "kernel": f"tcgen05 GEMM kernels: hgr::gemm_kernel<BN> / conv3x3_halo_kernel (implicit GEMM) and the chained "
                              f"stem_umma_kernel / gelan_tail_kernel / vit_block_kernel, {len(gem)} launches per step (the fused tcgen05 attention and "
                              f"pose-head kernels are listed in per_kernel against their own bounds)",
                    "share_of_step": g_ms / step_ms, "launch_ms_avg": g_ms / len(gem),
                    "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "flops_per_launch_avg": g_fl / len(gem)}
        # every launch against its own bound: tensor launches vs the sustained bf16 peak, the rest vs the copy bandwidth
        roofline["per_kernel"] = [
            {"launch": r["launch"], "ms": round(r["ms"], 4), "bound": "tensor" if r["kind"] != "memory" and r["tflops"]
             and r["tflops"] / pk["bf16_sustained"] >= (r["gbs"] or 0) / pk["hbm_gbs"] else "hbm",
             "tflops": round(r["tflops"], 1) if r["tflops"] else None, "gbs": round(r["gbs"], 1) if r["gbs"] else None,
             "frac": round(max((r["tflops"] or 0) / pk["bf16_sustained"], (r["gbs"] or 0) / pk["hbm_gbs"]), 3)}
            for r in table]
        if args.profile_out:
            Path(args.profile_out).write_text(json.dumps({"batch": B, "size": S, "step_ms_sum": step_ms,
                                                          "launches": table}, indent=1))

    # ---- memory-bound tail, stock-PyTorch GPU baseline, training step at the same N ----------
    mem_tail = eager = None
    if extras:
        conv1_row = next((r for r in table if r["launch"] == "encoder.conv1"), None)
        mem_tail = memory_tail_roofline(model, min(B, 1024), S, dev, pk["hbm_gbs"], conv1_row if B <= 1024 else None)
        eager = gpu_eager_rate(S, min(B, 1024), dev)
    train = other = None
    if not args.no_extras:
        del out
        torch.cuda.empty_cache()
        train = train_record(S, dev, rank, world, local_rank, barrier)
    if extras and S == 192 and B <= 1024:
        del model
        torch.cuda.empty_cache()
        other = forward_other_size(256, B, dev, pk)

    # ---- CPU baseline (bounded sample) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, kind, sample = cpu_forward_rate(S)
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        gf = GFLOP_PER_IMG.get(S)
        line = {
            "metric": "hand-crop images/s, MultiTaskNet forward", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload(B, S) if not strong else
                       (f"MultiTaskNet forward bf16, batch {args.total_batch} sharded across {world} B200 (no collective), "
                        f"3x{S}x{S} (BASELINE.json configs[2]); attention map not materialised"),
                       "batch_per_gpu": B, "image_size": S, "weights": "synthetic He-scaled, random BN statistics",
                       "l2": f"input batch is {x.numel() * 2 / 1e6:.0f} MB and every activation buffer of the step "
                             "is larger than the 126 MB L2, so no flush between iterations",
                       "net_gflop_per_image": gf,
                       "net_tensor_frac_of_sustained": value / world * gf * 1e9 / (pk["bf16_sustained"] * 1e12) if gf else None,
                       "net_tensor_frac_of_burst": value / world * gf * 1e9 / (pk["bf16_burst"] * 1e12) if gf else None,
                       "target_images_per_s_per_gpu": 0.5 * pk["bf16_burst"] * 1e12 / (gf * 1e9) if gf else None},
            "clocks": clocks, "gpu_launches": launches_per_step * K,
            "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "roofline_memory": mem_tail, "gpu_eager_baseline": eager, "train": train, "forward_256": other,
        }
        if e2e is not None:
            line["gpu_launches_e2e"] = launches_e2e * K
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
