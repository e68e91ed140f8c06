"""Host-side logic that needs no GPU: C-ABI export, parameter layout, packing, module contract, sharding."""
import os
import re
import socket
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import multitasknet_oracle as O

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from hgr_b200 import _lib
    header = (ROOT / "include" / "hgr_b200.h").read_text()
    declared = set(re.findall(r"HGR_API[^;(]*?\b(hgr_\w+)\s*\(", header))
    assert len(declared) >= 20
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hgr_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes prototypes and header disagree"
    assert lib.hgr_version() >= 100


def test_param_layout_is_dense_and_complete():
    from hgr_b200 import _lib, packing
    lay = _lib.param_layout(192, 21, 19)
    total = _lib.load().hgr_param_bytes(192, 21, 19)
    end = 0
    for name, off, nbytes, dt, dims in lay:
        assert off >= end and off % 1024 == 0, name
        assert nbytes == int(np.prod(dims)) * (4 if dt == _lib.F32 else 2), name
        end = off + nbytes
    assert end <= total
    sd = O.synthetic_state_dict(0)
    tensors = packing.packed_tensors(sd, 192, "cpu")
    assert {l[0] for l in lay} == set(tensors)
    block = packing.pack(sd, 192, 21, 19, torch.device("cpu"))
    assert block.numel() == total
    # workspace grows linearly with the batch and is non-trivial
    w1, w8 = _lib.load().hgr_workspace_bytes(192, 1), _lib.load().hgr_workspace_bytes(192, 8)
    assert w1 > 5_000_000 and 7.5 * w1 < w8 <= 8 * w1
    assert _lib.load().hgr_workspace_bytes(100, 1) == 0  # unsupported size is refused
    assert b"multiple of 64" in _lib.load().hgr_last_error()


def test_bn_folding_and_weight_order():
    """Packed [Cout][kh][kw][Cin] weights + scale/shift reproduce conv+BN of the oracle (fp32 on the packed values)."""
    from hgr_b200 import packing
    sd = O.synthetic_state_dict(3)
    t = packing.packed_tensors(sd, 192, "cpu")
    x = torch.randn(2, 64, 10, 10)
    p = "encoder.cspelan1.cv2.0.cv1"
    w = t[p + ".w"].float().permute(0, 3, 1, 2)  # back to (Cout, Cin, kh, kw)
    y = F.conv2d(x, w, None, 1, 1) * t[p + ".scale"].view(1, -1, 1, 1) + t[p + ".shift"].view(1, -1, 1, 1)
    sd_r = dict(sd)
    sd_r[p + ".conv.weight"] = sd[p + ".conv.weight"].to(torch.bfloat16).float()
    ref = O.conv_bn_act(sd_r, p, x, 3, 1, act=False)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)
    # conv1: scale folded into the weights, K padded 27 -> 32, k = (kh*3+kw)*3 + c
    w1 = t["encoder.conv1.w"].float()
    assert w1.shape == (64, 32) and torch.all(w1[:, 27:] == 0)
    scale, shift = packing.fold_bn(sd, "encoder.conv1")
    ref1 = (sd["encoder.conv1.conv.weight"] * scale.view(-1, 1, 1, 1))[5, 2, 1, 0]  # co=5, c=2, kh=1, kw=0
    assert abs(float(w1[5, (1 * 3 + 0) * 3 + 2]) - float(ref1)) <= abs(float(ref1)) * 2 ** -8
    torch.testing.assert_close(t["encoder.conv1.shift"], shift)
    torch.testing.assert_close(packing.sincos_table(12, 12), O.pos_emb_sincos_2d(12, 12, 256))
    assert t["decoder.pos_embedding"].shape == (144, 256)


def test_module_tree_and_contract():
    from hgr_b200 import MultiTaskNet
    m = MultiTaskNet(21, 19, [192, 192])
    sd = m.state_dict()
    spec = O.state_dict_spec()
    assert list(sd.keys()) == [k for k, _ in spec]
    for k, shape in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    assert sd["encoder.conv1.bn.num_batches_tracked"].dtype == torch.int64
    assert "decoder.pos_embedding" not in sd and m.decoder.pos_embedding.shape == (144, 256)
    m.load_state_dict(O.synthetic_state_dict(1), strict=True)
    # Lightning checkpoints prefix keys with "model." and export.py strips it (reference export.py:37-40)
    ckpt = {"model." + k: v for k, v in sd.items()}
    m.load_state_dict({k[len("model."):]: v for k, v in ckpt.items()}, strict=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.eval()(torch.zeros(1, 3, 192, 192))
    with pytest.raises(TypeError):
        m("not a tensor")
    with pytest.raises(ValueError):
        MultiTaskNet(21, 19, [192, 256])
    with pytest.raises(RuntimeError, match="parameter container"):
        m.encoder(torch.zeros(1, 3, 192, 192))
    assert MultiTaskNet(21, 19, [256, 256]).decoder.pos_embedding.shape == (256, 256)


def test_res_bottleneck_container_and_large_refusal():
    """model/gelan.py:90-121 is defined but never instantiated by gelan_spec ('small' only): the drop-in keeps the
    class with the reference's constructor and parameter tree, and refuses the 'large' backbone explicitly."""
    from hgr_b200.model import GELANNet, ResBottleneck
    blk = ResBottleneck(64, 64)
    keys = list(blk.state_dict().keys())
    want = [f"cv{i}.{leaf}" for i in (1, 2, 3)
            for leaf in ("conv.weight", "bn.weight", "bn.bias", "bn.running_mean", "bn.running_var",
                         "bn.num_batches_tracked")]
    assert keys == want
    assert tuple(blk.cv1.conv.weight.shape) == (32, 64, 1, 1) and tuple(blk.cv2.conv.weight.shape) == (32, 32, 3, 3)
    assert tuple(blk.cv3.conv.weight.shape) == (64, 32, 1, 1) and blk.add and blk.downsample is None
    assert not ResBottleneck(64, 128).add and not ResBottleneck(64, 64, shortcut=False).add
    with pytest.raises(RuntimeError, match="parameter container"):
        blk(torch.zeros(1, 64, 8, 8))
    with pytest.raises(ValueError, match="small"):
        GELANNet("large")


def test_signature_sees_every_kind_of_parameter_change():
    """The weight pack (and every CUDA graph / pipeline plan built on it) is keyed on this signature."""
    from hgr_b200 import MultiTaskNet
    m = MultiTaskNet(21, 19, [192, 192])
    a = m._signature()
    assert len(a) == 180 and m._signature() == a
    with torch.no_grad():
        m.encoder.cspelan2.cv4.bn.running_var.add_(1.0)          # buffer written in place
    b = m._signature()
    assert b != a
    m.load_state_dict(O.synthetic_state_dict(3), strict=True)    # parameters copied in place
    c = m._signature()
    assert c != b
    m.proj.weight = torch.nn.Parameter(torch.zeros_like(m.proj.weight))  # parameter object replaced
    d = m._signature()
    assert d != c
    m.decoder.cls_token.data = m.decoder.cls_token.data.clone()  # storage re-homed (what .to() does)
    assert m._signature() != d


def test_ops_refuse_cpu_inputs():
    from hgr_b200 import crop_normalize, get_max_preds
    with pytest.raises(RuntimeError):
        get_max_preds(torch.zeros(1, 1, 4, 4))
    with pytest.raises(AssertionError):
        get_max_preds(torch.zeros(1, 4, 4))
    with pytest.raises(RuntimeError):
        crop_normalize(torch.zeros(4, 4, 3, dtype=torch.uint8))


def test_shard_range_partitions_exactly():
    from hgr_b200.sharding import shard_range
    for total in [0, 1, 7, 8, 1024, 8192, 8191]:
        for world in [1, 2, 3, 4, 8]:
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from hgr_b200.sharding import max_over_ranks, shard_range, sum_over_ranks
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(101, rank, world)
    dist.barrier()
    slowest = max_over_ranks(10.0 + rank)
    units = sum_over_ranks(float(hi - lo))
    q.put((rank, lo, hi, slowest, units))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    """The N>1 plumbing of bench.py (shard, barrier, max-over-ranks time, summed units) on 2 CPU ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 51), (51, 101)]
    assert all(r[3] == 11.0 and r[4] == 101.0 for r in res)


def test_committed_ncu_capture_feeds_the_roofline_traffic():
    """bench.py takes roofline.traffic from the newest profiles/*_step_ncu_full.csv: the committed capture must parse,
    cover every tcgen05 launch of a forward (implicit-GEMM, halo, fused stem, fused GELAN tail and chained ViT kernels) and give a
    DRAM figure per launch of the order of the algorithmic bytes (hundreds of MB at batch 1024)."""
    import csv
    import importlib.util
    root = Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("bench_for_test", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    per_launch, name = bench.ncu_traffic_per_launch()
    assert name is not None and (root / "profiles" / name).exists()
    assert 1e8 < per_launch < 2e9, per_launch
    rows = list(csv.DictReader((root / "profiles" / name).open()))
    kernels = " ".join(r["Kernel Name"] for r in rows)
    for needle in ("gemm_kernel", "halo", "stem_umma_kernel", "gelan_tail_kernel", "attention_tc_kernel", "pose_head_tc_kernel"):
        assert needle in kernels, needle
    shares = [float(r["share_of_step_ncu"]) for r in rows]
    assert abs(sum(shares) - 1.0) < 1e-2


def test_numa_binding_helper_parses_cpulists_and_never_fails():
    """sharding.bind_host_to_device_node narrows the process's CPU affinity to the GPU's socket when sysfs says which
    one that is; without a GPU (here) or without NUMA information it must leave the affinity alone and return None."""
    import os
    from hgr_b200 import sharding
    assert sharding._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharding._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert sharding.bind_host_to_device_node(0) is None or os.sched_getaffinity(0) <= before
    if not __import__("torch").cuda.is_available():
        assert os.sched_getaffinity(0) == before
