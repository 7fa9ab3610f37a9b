import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # The shared libraries are git-ignored build products.  Build whatever is missing (g++/nvcc
    # only, no GPU needed) so a fresh checkout can run the suite.
    need = [os.path.join(ROOT, "vecchio_b200", "lib", "libvecchio_host.so"),
            os.path.join(ROOT, "vecchio_b200", "lib", "libvecchio_gpu.so"),
            os.path.join(ROOT, "vecchio_b200", "lib", "vecchio_gpu_render"),
            os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-j4", "all"], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def vb():
    import vecchio_b200
    return vecchio_b200


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def ctx(vb):
    """One GPU context for the whole session (gpu-marked tests only)."""
    c = vb.Context(0)
    yield c
    c.close()


_SCENES = {}


def get_scene(vb, name, seed=1, param=0):
    key = (name, seed, param)
    if key not in _SCENES:
        s = vb.Scene(name, seed=seed, param=param)
        _SCENES[key] = (s, s.next_camera())
    return _SCENES[key]
