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
