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


def test_entry_points_reject_bad_arguments_on_the_host():
    """Every argument check runs before any CUDA call, so the error convention (negative code + mgf_last_error text) is
    testable without a GPU: null tensors, empty batches and unsupported channel counts."""
    L = _lib.lib()
    P = ctypes.c_void_p(16)                                      # never dereferenced: the checks fail first
    cases = [
        ("mgf_vgg_conv1_fwd", (None, None, None, None, None, None, 1, 64, None), -1, b"null"),
        ("mgf_vgg_conv1_fwd", (P, None, None, P, P, P, 0, 64, None), -3, b"empty"),
        ("mgf_vgg_conv1_bwd", (None, P, None, None, 0.0, P, 1, 64, None), -1, b"null"),
        ("mgf_mapping_fwd", (None, None, None, None, 1, 11, None), -1, b"null"),
        ("mgf_mapping_fwd", (P, P, P, P, 0, 11, None), -3, b"empty"),
        ("mgf_mapping_bwd", (P, P, P, None, P, 1, 11, None), -1, b"null"),
        ("mgf_attn_fwd", (None,) * 9 + (1.0, 0.2, None, None, None, None, None, 1, 16, 32, 0, None), -1, b"null"),
        ("mgf_attn_fwd", (P,) * 9 + (1.0, 0.2, P, None, None, None, None, 1, 16, 48, 0, None), -3, b"C=48"),
        ("mgf_attn_bwd", (P,) * 10 + (1.0, 0.2, P, P, P, None, None, None, 0, 16, 32, 0, None), -3, b"empty"),
        ("mgf_lpips_head", (1, None, None, None, None, None, None, 0, 1, 16, 64, None), -1, b"null"),
        ("mgf_lpips_head", (1, P, P, P, None, None, P, 0, 1, 16, 96, None), -3, b"C=96"),
        ("mgf_upfir2_add", (P, None, P, (ctypes.c_float * 4)(1, 3, 3, 1), 1.0, 1, 4, 4, 24, None), -3, b"power of two"),
        ("mgf_maxpool2_fwd", (P, P, 1, 5, 4, 8, None), -3, b"maxpool2_fwd"),
    ]
    for name, args, code, text in cases:
        rc = getattr(L, name)(*args)
        assert rc == code, (name, rc, L.mgf_last_error())
        assert text in L.mgf_last_error(), (name, L.mgf_last_error())
    assert L.mgf_mapping_param_floats() == 4 * (6 * 1024 + 2 * 512 + 4 * 32) + 1056 + 4 * (2 * 1024 + 64) + 1056
