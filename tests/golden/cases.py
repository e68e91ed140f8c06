"""Seeded inputs shared by make_golden.py (which runs the reference on them) and the tests."""
import numpy as np


def heatmap_cases():
    """Crafted get_max_preds inputs: ties, all-negative, zeros, NaN, +-inf, 48x48, 64x64, non-square."""
    rng = np.random.default_rng(0)
    maps = rng.standard_normal((3, 21, 48, 48)).astype(np.float32)
    maps[0, 0] = -np.abs(maps[0, 0])            # all negative -> (0, 0)
    maps[0, 1] = 0.0                            # all zero -> (0, 0), maxval 0
    maps[0, 2] = 0.25                           # all tied positive -> index 0
    maps[0, 3, 10, 7] = 9.0
    maps[0, 3, 30, 40] = 9.0                    # two-way tie -> first
    maps[0, 4, 5, 5] = np.nan                   # NaN wins, masked to (0, 0)
    maps[0, 5, 47, 47] = 100.0                  # last element
    maps[0, 6, 0, 0] = 100.0                    # first element
    maps[0, 7] = -0.0
    maps[0, 8, 3, 3] = np.inf
    maps[0, 9] = -np.inf
    maps[0, 10, 1, 1] = np.nan
    maps[0, 10, 40, 2] = np.nan                 # two NaNs -> first
    maps64 = rng.standard_normal((2, 21, 64, 64)).astype(np.float32)
    maps64[1, 0, 63, 0] = 50.0
    rect = rng.standard_normal((2, 3, 5, 9)).astype(np.float32)   # non-square, odd sizes
    return {"maps48": maps, "maps64": maps64, "rect": rect}


def crop_image():
    """(16, 16, 3) uint8 holding every byte value in every channel."""
    img = np.stack([np.arange(256, dtype=np.uint8).reshape(16, 16)] * 3, axis=-1)
    img[..., 1] = img[..., 1][::-1]
    img[..., 2] = np.roll(img[..., 2], 7)
    return img


def crop_frame():
    """A 360 x 480 BGR uint8 frame: smooth gradients plus seeded noise (every tap of the bilinear stencil matters)."""
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:360, 0:480]
    base = np.stack([(xx * 255 // 479), (yy * 255 // 359), ((xx + yy) * 255 // 838)], axis=-1)
    noise = rng.integers(-40, 41, base.shape)
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def crop_boxes():
    """(detector box (x1, y1, x2, y2), crop size): inside, touching and crossing the frame border, tiny, whole frame."""
    return [((100, 80, 300, 290), 192), ((0, 0, 200, 150), 192), ((380, 250, 560, 430), 192),
            ((-40, -30, 90, 100), 192), ((210, 150, 240, 185), 192), ((5, 5, 475, 355), 192),
            ((120, 60, 360, 330), 256), ((33, 47, 161, 200), 64)]


def metric_cases():
    """(predicted heatmaps, ground-truth heatmaps) pairs for pose_accuracy: near-miss predictions, exact hits,
    invisible joints (target peak at x <= 1 or y <= 1), an all-invalid joint, 48x48 and 64x64."""
    rng = np.random.default_rng(21)
    out = {}
    for name, (b, j, h) in {"m48": (16, 21, 48), "m64": (5, 21, 64), "tiny": (2, 3, 8)}.items():
        tgt = np.zeros((b, j, h, h), dtype=np.float32)
        prd = rng.standard_normal((b, j, h, h)).astype(np.float32) * 0.05
        for n in range(b):
            for c in range(j):
                ty, tx = rng.integers(0, h, 2)
                if c == 0:
                    ty, tx = 0, int(rng.integers(0, h))  # never valid: y <= 1
                tgt[n, c, ty, tx] = 1.0
                dy, dx = rng.integers(-4, 5, 2)
                py, px = np.clip(ty + dy, 0, h - 1), np.clip(tx + dx, 0, h - 1)
                prd[n, c, py, px] = 1.0 + rng.random()
        out[name] = (prd, tgt)
    return out
