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
  roofline  the tcgen05 implicit-GEMM kernel (38 of the 54 launches of a step):
            algorithmic FLOPs / CUDA-event time of those launches, measured live
            in a profiling pass, against MEASURED_PEAKS.json
  cpu_baseline  the oracle (fp32 restatement of the reference forward) timed on
            the box's host cores on a bounded sample (rank 0, N = 1 only)

`--impl reference` times that CPU path alone (the reference is pure PyTorch on
CPU/cuDNN and /root/reference does not exist on the GPU box; the oracle port is
the reference's algorithm on the same ATen operators).
Under torchrun every rank drives one GPU with its own batch (weak scaling, no
data-path collective); rank 0 prints ONE JSON line.
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
    ap.add_argument("--no-e2e", action="store_true")
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
    rows = [r for r in csv.DictReader(files[-1].open()) if any(k in r["Kernel Name"] for k in ("gemm_kernel", "halo", "vit_block"))]
    if not rows:
        return None, None
    rd = [k for k in rows[0] if k.startswith("dram__bytes_read.sum")][0]
    wr = [k for k in rows[0] if k.startswith("dram__bytes_write.sum")][0]
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[rd.split("[")[1].rstrip("]")]
    total = sum(float(r[rd]) + float(r[wr]) for r in rows) * unit
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


def cpu_forward_rate(size, batch=32, budget_s=12.0, min_iters=3):
    """The oracle's fp32 forward on all host cores: (img/s, cores, description)."""
    import torch
    from oracle import multitasknet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.synthetic_state_dict(0)
    x = O.synthetic_images(batch, size, 1)
    O.multitasknet_forward(sd, x)  # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_iters or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        O.multitasknet_forward(sd, x)
        times.append(time.perf_counter() - t0)
        if len(times) >= 64:
            break
    med = statistics.median(times)
    return batch / med, cores, f"oracle fp32 forward, batch {batch} x {len(times)} iterations at {size}x{size}, median"


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
    sd = O.synthetic_state_dict(0)
    x = O.synthetic_images(batch, args.size, 1)
    for _ in range(max(1, min(args.warmup, 3))):
        O.multitasknet_forward(sd, x)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.multitasknet_forward(sd, x)
    dt = time.perf_counter() - t0
    v = batch * steps / dt
    sample = f"oracle fp32 forward (reference algorithm, torch CPU), batch {batch} x {steps} steps at {args.size}x{args.size}"
    print(json.dumps({
        "impl": "reference", "metric": "hand-crop images/s, MultiTaskNet forward", "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.batch, args.size), "batch_per_gpu": args.batch, "image_size": args.size,
                   "sample": f"each step is a bounded sample of the workload: {batch} of the {args.batch} crops, "
                             "fp32 on the host cores (the reference's own CPU path, BASELINE.json configs[0])"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
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
    tr = DataParallelTrainer(model, lr=1e-4)
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
                                   "block (BASELINE.json configs[4])",
                       "batch_per_gpu": B, "image_size": S,
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
    from hgr_b200.sharding import max_over_ranks

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MultiTaskNet path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    B, S, K, W = args.batch, args.size, args.steps, max(args.warmup, 3)
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
    value = world * B * K / (ms_total * 1e-3)
    assert torch.isfinite(out[0].float()).all() and torch.isfinite(out[1].float()).all(), "non-finite outputs"
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
        e2e = {"value": world * B * K / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": ms_e2e / K,
               "path": "HandPipeline: pinned uint8 crops -> H2D -> crop_normalize -> forward -> get_max_preds -> "
                       "logits+keypoints D2H, 2 lanes"}
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
            table.append({"launch": name, "kind": ["tcgen05_gemm", "mma_sync", "memory"][kind], "ms": t,
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
                              f"vit_block_kernel, {len(gem)} launches per step",
                    "share_of_step": g_ms / step_ms, "launch_ms_avg": g_ms / len(gem),
                    "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "flops_per_launch_avg": g_fl / len(gem)}
        if args.profile_out:
            Path(args.profile_out).write_text(json.dumps({"batch": B, "size": S, "step_ms_sum": step_ms,
                                                          "launches": table}, indent=1))

    # ---- CPU baseline (bounded sample) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_forward_rate(S)
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        gf = GFLOP_PER_IMG.get(S)
        line = {
            "metric": "hand-crop images/s, MultiTaskNet forward", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload(B, S),
                       "batch_per_gpu": B, "image_size": S, "weights": "synthetic He-scaled, random BN statistics",
                       "l2": f"input batch is {x.numel() * 2 / 1e6:.0f} MB and every activation buffer of the step "
                             "is larger than the 126 MB L2, so no flush between iterations",
                       "net_gflop_per_image": gf,
                       "net_tensor_frac_of_sustained": value / world * gf * 1e9 / (pk["bf16_sustained"] * 1e12) if gf else None,
                       "net_tensor_frac_of_burst": value / world * gf * 1e9 / (pk["bf16_burst"] * 1e12) if gf else None},
            "clocks": clocks, "gpu_launches": launches_per_step * K,
            "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        }
        if e2e is not None:
            line["gpu_launches_e2e"] = launches_e2e * K
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
