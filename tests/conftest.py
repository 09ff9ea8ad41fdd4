import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_image(g):
    """Regenerate a golden case's input with the oracle's generator and check its hash."""
    import hashlib
    from oracle import oracle as O
    gen = str(g["gen"])
    img = (O.blocks_v1 if gen == "blocks" else O.uniform_v1)(int(g["w"]), int(g["h"]), int(g["seed"]), int(g["frame"]))
    assert hashlib.sha256(img.tobytes()).hexdigest() == str(g["img_sha"])
    return img


EXTRACT_CASES = ["cfg1_752x480_1000", "cfg2_752x480_1200_f3", "kitti_1241x376_2000", "hd_1280x720_1000",
                 "uniform_400x300_500", "mono_lap_640x480_1000", "fisheye_lap_640x480_800_l6"]


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
