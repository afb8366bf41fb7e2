"""Pins oracle/resize.py (A9) bit-exactly against Pillow, the reference's dependency
(src/extractor/visualise_resnet.py:40-47, src/extractor/visualise_vit_layer.py:466-470)."""
import numpy as np
import pytest
from PIL import Image

from oracle import resize as R


@pytest.mark.parametrize("wh", [(960, 540), (1920, 1080), (404, 720), (224, 224), (300, 224), (224, 150), (97, 61)])
@pytest.mark.parametrize("filt", [R.BILINEAR, R.LANCZOS])
def test_resize_vs_pillow(wh, filt):
    rng = np.random.default_rng(wh[0] * 7 + filt)
    img = rng.integers(0, 256, (wh[1], wh[0], 3), dtype=np.uint8)
    pil = Image.BILINEAR if filt == R.BILINEAR else Image.LANCZOS
    ref = np.asarray(Image.fromarray(img).resize((224, 224), pil))
    assert np.array_equal(R.resize(img, 224, 224, filt), ref)


def test_torchvision_resize_is_pil_bilinear():
    from torchvision import transforms
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (270, 480, 3), dtype=np.uint8)
    ref = np.asarray(transforms.Resize((224, 224))(Image.fromarray(img)))
    assert np.array_equal(R.resize(img, 224, 224, R.BILINEAR), ref)
