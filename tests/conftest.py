import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        n_gpu = torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        n_gpu = 0
    for item in items:
        if "gpu" in item.keywords and n_gpu == 0:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "multigpu" in item.keywords and n_gpu < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 CUDA devices"))


# optional function-level coverage (tests/_cov.py): TDP_COV_DIR=<dir> python -m pytest ...
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _cov  # noqa: E402

_cov.start()


def pytest_sessionfinish(session, exitstatus):
    _cov.dump()
