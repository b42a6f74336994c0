import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def seq_small():
    """320x240 synthetic stream (same generator as the 640x480 bench stream)."""
    from lsd_slam_b200 import synth
    return synth.Sequence(320, 240, seed=1234)


@pytest.fixture(scope="session")
def frames_small(seq_small):
    """frame index -> (uint8 image, float32 z-depth); rendered lazily and cached."""
    cache = {}

    class _F:
        def __getitem__(self, k):
            if k not in cache:
                cache[k] = seq_small.render(k)
            return cache[k]
    return _F()


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.set_globals()          # reference defaults (settings.cpp:77-88), scalar path, 4 mapping threads
    return pyoracle


@pytest.fixture()
def gpu_ctx_small(seq_small):
    from lsd_slam_b200 import abi
    ctx = abi.Context(seq_small.w, seq_small.h, seq_small.K, device=0, max_frames=12)
    yield ctx
    ctx.close()
