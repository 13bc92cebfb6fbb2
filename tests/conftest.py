import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name has a dash -> importlib)."""
    return importlib.import_module("dreamerv3-torch_b200")


@pytest.fixture(scope="session")
def device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # the product has no fallback: the library must be the thing that runs
    lib = os.path.join(ROOT, "dreamerv3-torch_b200", "libdv3_b200.so")
    assert os.path.isfile(lib), "libdv3_b200.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    return "cuda:0"


@pytest.fixture
def knob(pkg):
    """Set a library environment knob (DV3_*) for one test: the library caches its environment,
    so every change is followed by dv3_reload_env(); restored afterwards."""
    import os
    saved = {}

    def set_(name, value):
        saved.setdefault(name, os.environ.get(name))
        os.environ[name] = value
        pkg._lib.lib().dv3_reload_env()

    yield set_
    for name, old in saved.items():
        if old is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = old
    pkg._lib.lib().dv3_reload_env()
