import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# the reference's pics.txt pairs (stereo_matching/pics.txt:1-10) + the classic 16-disparity tsukuba pair
PAIRS = {
    "tsukuba": ("im1.png", "im5.png"),
    "art": ("view1.png", "view5.png"),
    "teddy": ("im2.png", "im6.png"),
    "cones": ("im2.png", "im6.png"),
    "laundry": ("view1.png", "view5.png"),
    "sukub": ("imL.png", "imP.png"),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_rgba(path: str) -> np.ndarray:
    """8-bit PNG -> RGBA8 (H, W, 4), alpha 255 -- what lodepng::decode hands the reference (main.cpp:184-186)."""
    from PIL import Image
    return np.ascontiguousarray(np.array(Image.open(path).convert("RGBA"), dtype=np.uint8))


def load_pair(ds: str):
    l, r = PAIRS[ds]
    return load_rgba(os.path.join(GOLDEN, ds, l)), load_rgba(os.path.join(GOLDEN, ds, r))


def crop_pair(ds: str, x0: int, y0: int, w: int, h: int):
    L, R = load_pair(ds)
    return np.ascontiguousarray(L[y0:y0 + h, x0:x0 + w]), np.ascontiguousarray(R[y0:y0 + h, x0:x0 + w])


@pytest.fixture(scope="session")
def oracle():
    from oracle import asw_oracle
    asw_oracle.lib()
    return asw_oracle


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    """The CUDA context of the product library.  No fallback: fails loudly without a GPU."""
    from stereo_matchin_b200.api import AswContext
    c = AswContext(0)
    yield c
    c.close()
