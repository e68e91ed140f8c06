"""Bit-exact parity of the memory-bound tail: get_max_preds and crop normalisation.

Checked against (a) the golden vectors the REAL reference produced on the
crafted cases (tests/golden/make_golden.py) and (b) the oracle's numpy
restatement on seeded random maps at BASELINE.json's full batch size.
"""
import numpy as np
import pytest
import torch

from oracle import multitasknet_oracle as O
from tests.golden.cases import crop_image, heatmap_cases

pytestmark = pytest.mark.gpu
GOLD = __import__("pathlib").Path(__file__).resolve().parent / "golden"


def _same_bits(a: np.ndarray, b: np.ndarray):
    assert a.shape == b.shape and a.dtype == b.dtype == np.float32
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), \
        f"{np.sum(a.view(np.uint32) != b.view(np.uint32))} of {a.size} words differ"


@pytest.mark.parametrize("name", ["maps48", "maps64", "rect"])
def test_get_max_preds_golden(name):
    from hgr_b200 import get_max_preds
    gold = np.load(GOLD / "get_max_preds.npz")
    maps = heatmap_cases()[name]
    preds, maxvals = get_max_preds(torch.from_numpy(maps).cuda())
    _same_bits(preds.cpu().numpy(), gold["preds_" + name])
    _same_bits(maxvals.cpu().numpy(), gold["maxvals_" + name])
    # numpy in / numpy out, the reference's own signature
    p2, v2 = get_max_preds(maps)
    _same_bits(p2, gold["preds_" + name])
    _same_bits(v2, gold["maxvals_" + name])


@pytest.mark.parametrize("shape", [(1024, 21, 48, 48), (64, 21, 64, 64), (7, 5, 13, 11), (1, 1, 1, 1)])
def test_get_max_preds_random_vs_oracle(shape):
    from hgr_b200 import get_max_preds
    g = torch.Generator().manual_seed(shape[0])
    h = torch.randn(shape, generator=g)
    h[h.abs() < 0.05] = 0.0  # plenty of exact ties at zero
    preds, maxvals = get_max_preds(h.cuda())
    rp, rv = O.get_max_preds(h.numpy())
    _same_bits(preds.cpu().numpy(), rp)
    _same_bits(maxvals.cpu().numpy(), rv)


def test_get_max_preds_bf16_input_matches_float_of_same_values():
    from hgr_b200 import get_max_preds
    h = torch.randn(32, 21, 48, 48, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    preds, maxvals = get_max_preds(h.cuda())
    rp, rv = O.get_max_preds(h.float().numpy())
    _same_bits(preds.cpu().numpy(), rp)
    _same_bits(maxvals.cpu().numpy(), rv)


def test_get_max_preds_edge_shapes_and_errors():
    from hgr_b200 import get_max_preds
    p, v = get_max_preds(torch.empty(0, 21, 48, 48, device="cuda"))
    assert p.shape == (0, 21, 2) and v.shape == (0, 21, 1)
    with pytest.raises(AssertionError):
        get_max_preds(torch.zeros(21, 48, 48, device="cuda"))  # reference asserts ndim == 4
    with pytest.raises(AssertionError):
        get_max_preds([[1.0]])
    with pytest.raises(RuntimeError):
        get_max_preds(torch.zeros(1, 1, 4, 4))  # CPU tensor: no fallback


def test_crop_normalize_golden_and_oracle():
    from hgr_b200 import crop_normalize
    gold = np.load(GOLD / "crop_normalize.npz")["out"]
    img = crop_image()
    out = crop_normalize(torch.from_numpy(img).cuda())
    _same_bits(out.cpu().numpy(), gold)
    # a full-size batch against the oracle restatement
    rng = np.random.default_rng(1)
    batch = rng.integers(0, 256, size=(5, 192, 192, 3), dtype=np.uint8)
    got = crop_normalize(torch.from_numpy(batch).cuda()).cpu().numpy()
    ref = np.concatenate([O.crop_normalize(b) for b in batch], 0)
    _same_bits(got, ref)
    # odd sizes take the scalar path
    odd = rng.integers(0, 256, size=(2, 7, 5, 3), dtype=np.uint8)
    got = crop_normalize(torch.from_numpy(odd).cuda()).cpu().numpy()
    _same_bits(got, np.concatenate([O.crop_normalize(b) for b in odd], 0))
    with pytest.raises(RuntimeError):
        crop_normalize(torch.from_numpy(img))


# ------------------------------------------------------------------ crop front-end ----
def test_crop_warp_normalize_golden_is_bit_exact():
    """detect.py:92-117 fused on the device against cv2.warpAffine + the reference's get_affine_transform
    (golden CRC32 of the fp32 result) and against the oracle's fixed-point restatement, word for word."""
    import zlib
    from hgr_b200 import crop_warp_normalize
    from tests.golden.cases import crop_boxes, crop_frame
    gold = np.load(GOLD / "crop_warp.npz")
    frame = crop_frame()
    dframe = torch.from_numpy(frame).cuda()
    for i, (bbox, size) in enumerate(crop_boxes()):
        out = crop_warp_normalize(dframe, [bbox], size=size).cpu().numpy()
        ref, _ = O.process_image_for_classification(frame, bbox, size)
        _same_bits(out, ref)
        assert zlib.crc32(out.tobytes()) == int(gold[f"out_crc_{i}"][0]), i
        # explicit matrix path (the matrix cv2.getAffineTransform produced inside the reference)
        out2 = crop_warp_normalize(dframe, None, size=size, trans=[gold[f"trans_{i}"]]).cpu().numpy()
        _same_bits(out2, ref)


def test_crop_warp_normalize_batch_of_frames_and_bf16():
    from hgr_b200 import crop_warp_normalize
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (3, 120, 160, 3), dtype=np.uint8)
    boxes = [(10, 20, 90, 100), (-5, -5, 60, 70), (100, 60, 170, 130), (40, 30, 120, 110), (0, 0, 160, 120)]
    index = [0, 1, 2, 1, 0]
    out = crop_warp_normalize(torch.from_numpy(frames).cuda(), boxes, frame_index=index, size=64)
    outb = crop_warp_normalize(torch.from_numpy(frames).cuda(), boxes, frame_index=index, size=64, dtype=torch.bfloat16)
    for n, (b, f) in enumerate(zip(boxes, index)):
        ref, _ = O.process_image_for_classification(frames[f], b, 64)
        _same_bits(out[n: n + 1].cpu().numpy(), ref)
        assert torch.equal(outb[n].float().cpu(), torch.from_numpy(ref[0]).to(torch.bfloat16).float())
    assert crop_warp_normalize(torch.from_numpy(frames).cuda(), [], size=64).shape == (0, 3, 64, 64)
    with pytest.raises(ValueError):
        crop_warp_normalize(torch.from_numpy(frames).cuda(), boxes, frame_index=[0, 1, 2, 3, 0], size=64)
    with pytest.raises(RuntimeError):
        crop_warp_normalize(torch.from_numpy(frames), boxes, size=64)


# ------------------------------------------------------------------ metrics tail ----
def test_pose_accuracy_matches_reference_golden_exactly():
    """libs.metrics.pose_accuracy on the device (no heatmap D2H copy): float64 accuracies, count and decoded
    keypoints identical to the real reference's on the crafted cases and to the oracle on a full training batch."""
    from hgr_b200 import pose_accuracy
    from tests.golden.cases import metric_cases
    g = np.load(GOLD / "pose_accuracy.npz")
    for name, (prd, tgt) in metric_cases().items():
        acc, avg, cnt, pred = pose_accuracy(torch.from_numpy(prd).cuda(), torch.from_numpy(tgt).cuda())
        assert np.array_equal(acc.cpu().numpy(), g["acc_" + name]), name
        assert float(avg) == g["avg_" + name][0] and int(cnt) == g["cnt_" + name][0]
        _same_bits(pred.cpu().numpy(), g["pred_" + name])
    _, target, _ = O.synthetic_targets(32, 192, seed=9)
    out = target + 0.3 * torch.randn(target.shape, generator=torch.Generator().manual_seed(1))
    acc, avg, cnt, _ = pose_accuracy(out.cuda(), target.cuda())
    racc, ravg, rcnt, _ = O.pose_accuracy(out.numpy(), target.numpy())
    assert np.array_equal(acc.cpu().numpy(), racc) and float(avg) == ravg and int(cnt) == rcnt
    with pytest.raises(RuntimeError):
        pose_accuracy(out, target)
