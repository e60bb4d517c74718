"""CPU: the preprocessing oracle is pinned bit-for-bit against the Pillow / torchvision installed in the image, and the product's
host logic (coefficient tables, output-size and crop rules) agrees with both."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import preprocess as o_pre

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _product_module():
    sys.path.insert(0, ROOT)
    import patchioner_b200  # noqa: F401  (loads the package under its importable alias)
    from patchioner_b200 import preprocess

    return preprocess


@pytest.mark.parametrize("H,W,ow,oh", [(480, 640, 691, 518), (300, 200, 518, 777), (700, 1100, 352, 224), (224, 224, 224, 224), (37, 53, 224, 224)])
def test_oracle_resize_is_pillow_bit_exact(H, W, ow, oh):
    img = np.random.RandomState(H + W).randint(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(o_pre.pil_bicubic_resize(img, ow, oh), ref)


def test_product_tables_match_oracle_and_torchvision_sizes():
    pre = _product_module()
    for a, b in [(640, 691), (480, 518), (1500, 777), (53, 224), (518, 518)]:
        k, bounds, ksize = pre.resample_table(a, b)
        ko, bo = o_pre._coeffs(a, b)
        assert np.array_equal(k.astype(np.int64), ko) and np.array_equal(bounds.astype(np.int64), bo) and ksize == ko.shape[1]
    import torchvision.transforms as T

    for (w, h) in [(640, 480), (480, 640), (1001, 333), (518, 518), (519, 1000)]:
        im = Image.new("RGB", (w, h))
        for size in (224, 518):
            assert pre.resized_size(w, h, size, True) == T.Resize(size)(im).size
            assert pre.resized_size(w, h, size, False) == T.Resize((size, size))(im).size
