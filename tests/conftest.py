import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
  if p not in sys.path:
    sys.path.insert(0, p)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
  try:
    import ctypes
    from qcc_b200 import _cabi
    n = ctypes.c_int(0)
    return _cabi.lib().qb_device_count(ctypes.byref(n)) == 0 and n.value > 0
  except Exception:  # pylint: disable=broad-except
    return False


@pytest.fixture(scope="session")
def has_gpu():
  return _has_gpu()


def pytest_collection_modifyitems(config, items):
  # `-m gpu` on a box without a GPU must fail loudly, not skip: the product has no CPU path.
  # Without -m, GPU tests are skipped here in the CPU container.
  if config.getoption("-m"):
    return
  if _has_gpu():
    return
  skip = pytest.mark.skip(reason="no GPU in this container (run with -m gpu on the B200 box)")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)
