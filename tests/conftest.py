import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure).  Built on demand with gcc."""
    from oracle import orc as o
    o.build()
    tex = os.path.join(ROOT, "scenes", "txtrs_argb")
    if not os.path.isdir(tex) or len(os.listdir(tex)) < 8:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import decode_textures
        decode_textures.main(os.path.join(ROOT, "scenes", "txtrs"), tex)
    if not os.path.exists(os.path.join(ROOT, "scenes", "bun69k.cli")):
        import subprocess
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_bun69k.py")])
    return o


@pytest.fixture(scope="session")
def drt():
    import distraytracer_old_b200 as d
    from distraytracer_old_b200 import build
    build.build()
    d.load_library()
    return d


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))


@pytest.fixture(scope="session")
def gpu_ctx_factory(drt):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")

    def make(cols=300, rows=300, **kw):
        return drt.Context(device=0, cols=cols, rows=rows, **kw)
    return make
