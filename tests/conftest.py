import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")
    config.addinivalue_line("markers", "multigpu: spawns torch.distributed.run over >= 2 GPUs of the box (skips on fewer)")


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import poi_b200  # noqa: F401
    from poi_b200.engine import Engine
    return Engine.get(0)
