import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # plain `pytest tests` on a box without a GPU: the gpu-marked tests are skipped, not failed
    if os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0"):
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _oracles_built():
    # builds the C port always, and oracle/_ref when /root/reference is present (this container);
    # on the GPU box the prebuilt oracle/_ref/*.so that travelled with the snapshot is used as is.
    import oracle as ora
    ora.ensure_built()
    yield
