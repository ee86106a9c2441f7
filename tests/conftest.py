import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(autouse=True)
def _restore_forward_dtype(request):
    """GPU tests may switch the engine's 16-bit forward type; every test starts from and returns to the library default."""
    yield
    if "gpu" in request.keywords:
        import torch
        if torch.cuda.is_available():
            from morphganformer_b200 import _lib
            _lib.set_forward_dtype(_lib.DEFAULT_FORWARD_DTYPE)
