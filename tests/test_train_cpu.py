"""Training step, CPU side: the oracle's restatement of the reference's train-mode forward / losses / AdamW
against golden vectors produced by the REAL reference (tests/golden/make_golden_train.py), the flat
parameter layout of the C ABI, and the data-parallel gradient exchange over 2 gloo ranks."""
import os
import socket
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import multitasknet_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("size,seed,batch", [(64, 5, 4), (192, 11, 2)])
def test_oracle_train_step_matches_reference(size, seed, batch):
    g = np.load(GOLD / f"train_step_s{size}.npz")
    torch.set_num_threads(8)
    sd = O.synthetic_state_dict(seed)
    x = O.synthetic_images(batch, size, seed + 1)
    labels, target, weight = O.synthetic_targets(batch, size, seed=seed + 2)
    loss3, grads, stats, (cls, hm) = O.train_step_grads(sd, x, labels, target, weight)
    np.testing.assert_allclose(loss3.numpy(), g["loss3"], rtol=2e-5)
    np.testing.assert_allclose(cls.numpy(), g["logits"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(hm[:, :, ::4, ::4].numpy(), g["heat_sub"], rtol=1e-3, atol=1e-4)
    assert len(grads) == 114
    for k, gr in grads.items():
        ref_norm = float(g["gnorm/" + k][0])
        assert abs(float(gr.double().norm()) - ref_norm) <= 2e-3 * ref_norm + 1e-9, k
        sub = gr.flatten()[:: max(1, gr.numel() // 512)].numpy() if gr.numel() > 4096 else gr.numpy()
        ref = g["gsub/" + k]
        assert np.abs(sub - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-9, k
    for k, v in stats.items():
        np.testing.assert_allclose(v.numpy(), g["stat/" + k], rtol=1e-4, atol=1e-6, err_msg=k)
    # torch.optim.AdamW's first step, restated
    for k in ["encoder.conv2.conv.weight", "decoder.cls_token", "decoder.mlp_head.1.bias"]:
        p0 = sd[k].float()
        p1, _, _ = O.adamw_step(p0, grads[k], torch.zeros_like(p0), torch.zeros_like(p0), 1)
        ref = g["pnew_sum/" + k]
        assert abs(float(p1.double().sum()) - ref[0]) <= 1e-4 * ref[1] + 1e-6, k


def test_train_param_layout_follows_state_dict_order():
    from hgr_b200 import _lib
    lay = _lib.train_param_layout(21, 19)
    spec = [(k, s) for k, s in O.state_dict_spec() if "running_" not in k and not k.endswith("num_batches_tracked")]
    assert [l[0] for l in lay] == [k for k, _ in spec]
    end = 0
    for (name, off, n), (_, shape) in zip(lay, spec):
        assert off >= end and off % 64 == 0, name
        assert n == int(np.prod(shape)), name
        end = off + n
    assert end <= _lib.load().hgr_train_param_floats(21, 19)
    bn = _lib.train_bnstat_layout()
    assert len(bn) == 44 and bn[0][0] == "encoder.conv1.bn.running_mean" and bn[1][0] == "encoder.conv1.bn.running_var"
    assert _lib.load().hgr_train_workspace_bytes(192, 21, 19, 32) > 100_000_000
    assert _lib.load().hgr_train_workspace_bytes(192, 21, 19, 1) == 0  # batch statistics need batch >= 2


def test_gradient_buckets_follow_the_order_the_backward_completes_them():
    """grad_buckets: the three contiguous ranges of the flat gradient block that hgr_train_backward_part 0, 1, 2
    complete (heads + transformer + proj; down2 + cspelan3; conv1 .. cspelan2), in that order, covering every
    parameter exactly once - the ranges the data-parallel step may all-reduce while the next part runs."""
    from hgr_b200 import _lib
    from hgr_b200.training import grad_buckets
    lay = _lib.train_param_layout(21, 19)
    total = lay[-1][1] + lay[-1][2]
    b = grad_buckets(21, 19)
    assert len(b) == 3 and b[2][0] == 0 and b[0][1] == total
    assert b[2][1] == b[1][0] and b[1][1] == b[0][0]
    owner = {}
    for k, (lo, hi) in enumerate(b):
        for name, off, n in lay:
            if lo <= off < hi:
                assert off + n <= hi, name
                owner[name] = k
    assert len(owner) == len(lay)
    assert owner["proj.weight"] == 0 and owner["decoder.simple_decoder.1.bias"] == 0 and owner["decoder.cls_token"] == 0
    assert owner["encoder.down2.conv.weight"] == 1 and owner["encoder.cspelan3.cv4.bn.bias"] == 1
    assert owner["encoder.conv1.conv.weight"] == 2 and owner["encoder.cspelan2.cv4.bn.bias"] == 2


def test_train_mode_refuses_cpu():
    from hgr_b200 import MultiTaskNet
    m = MultiTaskNet(21, 19, [64, 64]).train()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(2, 3, 64, 64))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from hgr_b200.training import allreduce_sum_
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    local = torch.randn(4096, generator=g)  # this rank's flat gradient block
    flat = local.clone()
    w = allreduce_sum_(flat)
    # identical AdamW update on every rank with grad_scale = 1 / world
    p0 = torch.ones(4096)
    p1, _, _ = O.adamw_step(p0, flat / w, torch.zeros(4096), torch.zeros(4096), 1)
    q.put((rank, w, local.numpy(), flat.numpy(), p1.numpy()))  # by value: the sender may exit before the receiver reads
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_over_gloo():
    """The exchange step of the DP trainer (one SUM all-reduce over the flat gradient block, mean via
    grad_scale) on 2 CPU ranks: every rank ends with the mean of the per-rank gradients and the same update."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    total = res[0][2] + res[1][2]
    for r in res:
        assert r[1] == 2
        np.testing.assert_allclose(r[3], total, rtol=1e-6, atol=1e-6)
    assert np.array_equal(res[0][4], res[1][4])
