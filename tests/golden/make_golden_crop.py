"""Generates tests/golden/crop_warp.npz with the REAL reference pieces of detect.py:92-117:
libs.transforms.get_affine_transform (imported from /root/reference) + cv2.warpAffine(INTER_LINEAR) + the
normalisation arithmetic of detect.py:106-112 (detect.py itself imports onnxruntime, which is not installed, so
its method body is replayed line by line).  Run in the build container only.
"""
import sys
import zlib
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")
from libs.transforms import get_affine_transform  # noqa: E402  (the reference)

from tests.golden.cases import crop_boxes, crop_frame  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    frame = crop_frame()
    d = {}
    for i, (bbox, size) in enumerate(crop_boxes()):
        x1, y1, x2, y2 = bbox
        c = np.array([(x1 + x2) / 2, (y1 + y2) / 2], dtype=np.float32)
        origin_size = max(x2 - x1, y2 - y1) * 1.0
        trans = get_affine_transform(c, 1, 0, origin_size, [size, size])
        img = cv2.warpAffine(frame, trans, (int(size), int(size)), flags=cv2.INTER_LINEAR)
        im = img.transpose((2, 0, 1)).astype(np.float32)
        im /= 255
        mean = np.array([0.485, 0.456, 0.406], dtype=np.float32)
        std = np.array([0.229, 0.224, 0.225], dtype=np.float32)
        im = (im - mean.reshape(3, 1, 1)) / std.reshape(3, 1, 1)
        im = np.ascontiguousarray(np.expand_dims(im, 0))
        # the frame is noisy (incompressible): keep CRC32s of the full results plus a strided subsample
        d[f"trans_{i}"] = trans
        d[f"crop_crc_{i}"] = np.array([zlib.crc32(np.ascontiguousarray(img).tobytes())], dtype=np.uint32)
        d[f"out_crc_{i}"] = np.array([zlib.crc32(im.tobytes())], dtype=np.uint32)
        d[f"crop_sub_{i}"] = img[::4, ::4].copy()
    np.savez_compressed(OUT / "crop_warp.npz", **d)
    print("wrote", len(crop_boxes()), "cases")


if __name__ == "__main__":
    main()
