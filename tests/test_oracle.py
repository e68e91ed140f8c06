"""The oracle restatement against the golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU only; this is the pin that lets the GPU tests trust the oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import multitasknet_oracle as O
from tests.golden.cases import crop_image, heatmap_cases

GOLD = Path(__file__).resolve().parent / "golden"


def _stats(t):
    t = t.double()
    return np.array([t.mean().item(), t.std().item(), t.abs().sum().item()])


@pytest.mark.parametrize("size,seed,batch", [(192, 0, 4), (192, 7, 2), (256, 3, 2)])
def test_forward_matches_reference(size, seed, batch):
    gold = np.load(GOLD / f"multitasknet_s{size}_seed{seed}.npz")
    sd = O.synthetic_state_dict(seed)
    x = O.synthetic_images(batch, size, seed + 1)
    np.testing.assert_allclose(_stats(x), gold["x_stats"], rtol=1e-9)  # same seeded inputs as the generator saw
    taps = {}
    cls, hm, attn = O.multitasknet_forward(sd, x, taps)
    np.testing.assert_allclose(cls.numpy(), gold["logits"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(hm[:, :, ::4, ::4].numpy(), gold["heat_sub"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(attn[:, :, ::8, ::8].numpy(), gold["attn_sub"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(_stats(hm), gold["heat_stats"], rtol=1e-5)
    np.testing.assert_allclose(_stats(attn), gold["attn_stats"], rtol=1e-5)
    for name in ["a1", "a2", "o1", "d1", "o2", "d2", "o3", "proj"]:
        np.testing.assert_allclose(_stats(taps[name]), gold["stats_" + name], rtol=1e-5, err_msg=name)


def test_weight_recipe_is_not_degenerate():
    """with default init every image gives the same outputs; the synthetic recipe must not."""
    sd = O.synthetic_state_dict(0)
    x = O.synthetic_images(6, 192, 2)
    taps = {}
    cls, hm, _ = O.multitasknet_forward(sd, x, taps)
    assert taps["o3"].std() > 0.1
    assert cls.std(0).mean() > 0.05 and hm.std(0).mean() > 0.05
    assert len(set(cls.argmax(1).tolist())) >= 2


def test_state_dict_spec_has_the_reference_keys():
    spec = O.state_dict_spec()
    assert len(spec) == 180
    n_param = sum(int(np.prod(s)) for k, s in spec if "running" not in k and "num_batches" not in k)
    assert n_param == 7_409_000


def test_get_max_preds_matches_reference():
    gold = np.load(GOLD / "get_max_preds.npz")
    for name, maps in heatmap_cases().items():
        p, v = O.get_max_preds(maps)
        assert np.array_equal(p.view(np.uint32), gold["preds_" + name].view(np.uint32)), name
        assert np.array_equal(v.view(np.uint32), gold["maxvals_" + name].view(np.uint32)), name
    with pytest.raises(AssertionError):
        O.get_max_preds(np.zeros((2, 3, 4), np.float32))
    with pytest.raises(AssertionError):
        O.get_max_preds(torch.zeros(1, 1, 2, 2))


def test_crop_normalize_matches_reference():
    gold = np.load(GOLD / "crop_normalize.npz")["out"]
    out = O.crop_normalize(crop_image())
    assert out.dtype == np.float32 and np.array_equal(out.view(np.uint32), gold.view(np.uint32))
