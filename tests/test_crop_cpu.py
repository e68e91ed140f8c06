"""Crop front-end, CPU side: the oracle's fixed-point restatement of cv2.warpAffine(INTER_LINEAR)
and of libs/transforms.get_affine_transform against golden vectors made with the REAL cv2 + reference functions
(tests/golden/make_golden_crop.py), and the product's host-side matrix helpers against the oracle."""
import zlib
from pathlib import Path

import numpy as np

from oracle import multitasknet_oracle as O
from tests.golden.cases import crop_boxes, crop_frame

GOLD = Path(__file__).resolve().parent / "golden"


def test_oracle_warp_affine_is_bit_exact_against_cv2():
    g = np.load(GOLD / "crop_warp.npz")
    frame = crop_frame()
    for i, (bbox, size) in enumerate(crop_boxes()):
        out, trans = O.process_image_for_classification(frame, bbox, size)
        np.testing.assert_allclose(trans, g[f"trans_{i}"], rtol=0, atol=1e-12)
        crop = O.warp_affine_linear_u8(frame, g[f"trans_{i}"], size, size)
        assert np.array_equal(crop[::4, ::4], g[f"crop_sub_{i}"]), i
        assert zlib.crc32(np.ascontiguousarray(crop).tobytes()) == int(g[f"crop_crc_{i}"][0]), i
        assert out.shape == (1, 3, size, size) and out.dtype == np.float32
        assert zlib.crc32(out.tobytes()) == int(g[f"out_crc_{i}"][0]), i


def test_host_matrix_helpers_match_the_oracle():
    from hgr_b200.ops import box_to_affine, invert_affine
    for bbox, size in crop_boxes():
        x1, y1, x2, y2 = bbox
        c = np.array([(x1 + x2) / 2, (y1 + y2) / 2], dtype=np.float32)
        ref = O.get_affine_transform(c, 1, 0, max(x2 - x1, y2 - y1) * 1.0, [size, size])
        got = box_to_affine(bbox, size)
        assert np.array_equal(got, ref)
        assert np.array_equal(invert_affine(got), O.invert_affine(ref))


def test_oracle_pose_accuracy_matches_reference():
    """libs.metrics.pose_accuracy restated (oracle) against golden vectors from the real reference function."""
    from tests.golden.cases import metric_cases
    g = np.load(GOLD / "pose_accuracy.npz")
    for name, (prd, tgt) in metric_cases().items():
        acc, avg_acc, cnt, pred = O.pose_accuracy(prd, tgt)
        assert np.array_equal(acc, g["acc_" + name]) and avg_acc == g["avg_" + name][0] and cnt == g["cnt_" + name][0]
        assert np.array_equal(pred, g["pred_" + name])
