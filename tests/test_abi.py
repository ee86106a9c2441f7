"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/mgf.h declares."""
import ctypes
import os
import re
import pytest
from morphganformer_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mgf.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgf_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(L, n), "missing export %s" % n


def test_python_signatures_cover_header():
    _lib.register_all()
    assert sorted(_lib.SIGNATURES) == _declared()


def test_version_and_error_text():
    L = _lib.lib()
    assert L.mgf_version() >= 100
    # bad argument is rejected on the host before any CUDA call
    rc = L.mgf_bias_act(None, None, None, None, None, None, 0, 0, 3, 0.2, 1.0, -1.0, 16, 1, 1, None)
    assert rc < 0 and b"non-null" in L.mgf_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from morphganformer_b200.torch_utils.ops import bias_act, upfirdn2d
    with pytest.raises(RuntimeError):
        bias_act.bias_act(torch.zeros(4, 4), act="lrelu")
    with pytest.raises(RuntimeError):
        upfirdn2d.upfirdn2d(torch.zeros(1, 1, 4, 4), upfirdn2d.setup_filter([1, 3, 3, 1]))
    with pytest.raises(NotImplementedError):
        bias_act.bias_act(torch.zeros(4, 4), act="lrelu", impl="ref")
