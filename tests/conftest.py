import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# Measured parity figures (per-pixel / image relative L2, ray-count deltas ...) are collected here and printed in the
# terminal summary, so that a passing `pytest -q` run still shows the numbers next to the bounds they were checked against.
_MEASURED = []


def measured(line):
    _MEASURED.append(line)
    print(line)


def pytest_terminal_summary(terminalreporter):
    if _MEASURED:
        terminalreporter.write_sep("-", "measured parity figures (value <= asserted bound)")
        for line in _MEASURED:
            terminalreporter.write_line(line)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build every native component once (no-op when the .so files are current)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def akr():
    import akari_render_b200
    return akari_render_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding


@pytest.fixture(scope="session")
def tables(akr):
    return akr.sampler_tables()


CBOX = os.path.join(ROOT, "scenes", "cbox", "scene.json")
CBOX_PT = os.path.join(ROOT, "scenes", "cbox", "pt.json")


@pytest.fixture()
def cbox(akr):
    def make(width, height):
        return akr.load_scene(CBOX).set_resolution(width, height)
    return make


@pytest.fixture()
def cbox_task(akr):
    def make(spp, **kw):
        t = akr.RenderTask.from_file(CBOX_PT)
        t.pt.spp = spp
        for k, v in kw.items():
            setattr(t.pt, k, v)
        return t
    return make


def rel_l2_per_pixel(a, b):
    a = a.reshape(-1, 3).astype(np.float64)
    b = b.reshape(-1, 3).astype(np.float64)
    return np.linalg.norm(a - b, axis=1) / (np.linalg.norm(b, axis=1) + 1e-3)


def image_rel_l2(a, b):
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
